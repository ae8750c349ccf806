#!/bin/bash
# round-2 GPU call Q (1 GPU): the DMMA + DFMA hybrid restart GEMM — tests, standalone time, headline job
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_primitives.py -m gpu -q --timeout 300 -k "basistransform" > gpurun_out/r02q_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02q_pytest.log | tail -8
for h in 0 2; do
  B2K_TRANSFORM_HYB=$h timeout 200 python tools/microbench.py --reps 8 --k 60 2>&1 | grep basistransform | sed "s/^/hyb$h /" | cut -c1-150 | tee -a gpurun_out/r02q_transform.log
done
for h in 0 2; do
  B2K_TRANSFORM_HYB=$h timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --extra '' > gpurun_out/r02q_bench_hyb$h.json 2> gpurun_out/r02q_bench_hyb$h.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02q_bench_hyb$h.json').read().strip().splitlines()[-1])
    print('hyb$h', round(d['value'],1), 'it/s', {k:(v['avg_ms'],v['frac']) for k,v in d['kernels'].items()}, d['parity']['ok'], d['parity'].get('max_rel_diff_ritz'))
except Exception as e:
    print('hyb$h failed', e); print(open('gpurun_out/r02q_bench_hyb$h.err').read()[-1500:])
PY
done
