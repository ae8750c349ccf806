#!/bin/bash
# round-2 GPU call O (1 GPU): does the persisting-L2 set-aside slow plain streaming kernels?  Probe with and without it,
# CG breakdown with B2K_L2_CARVE=0, headline bench with B2K_L2_CARVE=0; block path after the PP templating
mkdir -p gpurun_out
( timeout 60 tools/probes/stream_probe 10000000 0; timeout 60 tools/probes/stream_probe 10000000 0 1000000000 ) 2>&1 | grep -E "set-aside|1r\+1rw U4 4|2r\+2rw U2 4|3r\+2rw U2 4|chunk   2r" | tee gpurun_out/r02o_probe.txt
echo "--- default carve"; timeout 200 python tools/cg_breakdown.py 2>&1 | tee gpurun_out/r02o_cg_default.txt
echo "--- B2K_L2_CARVE=0"; B2K_L2_CARVE=0 timeout 200 python tools/cg_breakdown.py 2>&1 | tee gpurun_out/r02o_cg_nocarve.txt
for c in default 0 33554432; do
  if [ $c = default ]; then unset B2K_L2_CARVE; else export B2K_L2_CARVE=$c; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --extra '' > gpurun_out/r02o_bench_$c.json 2> gpurun_out/r02o_bench_$c.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02o_bench_$c.json').read().strip().splitlines()[-1])
    print('carve=$c', round(d['value'],1), 'it/s', {k:(v['avg_ms'],v['frac']) for k,v in d['kernels'].items()}, d['parity']['ok'])
except Exception as e:
    print('carve=$c failed', e); print(open('gpurun_out/r02o_bench_$c.err').read()[-1500:])
PY
done
unset B2K_L2_CARVE
timeout 600 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_solvers.py -m gpu -q --timeout 300 -k "block or Block or spmm or toric" > gpurun_out/r02o_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02o_pytest.log | tail -8
timeout 120 python tools/block_probe.py fast 3 3 2>&1 | tail -2
