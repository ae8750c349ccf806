#!/bin/bash
# round-2 GPU call R (1 GPU): fused block sweep (tests, rate with and without), DFMA 8x9 restart GEMM (tests, standalone,
# in the job)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_solvers.py -m gpu -q --timeout 300 -k "block or Block or spmm or toric or basistransform" > gpurun_out/r02r_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02r_pytest.log | tail -8
timeout 120 python tools/block_probe.py fast 3 3 2>&1 | tail -2
B2K_BLOCK_FUSE=0 timeout 120 python tools/block_probe.py fast 3 3 2>&1 | tail -1 | sed 's/^/nofuse /'
for h in 0 2; do
  B2K_TRANSFORM_HYB=$h timeout 200 python tools/microbench.py --reps 8 --k 60 2>&1 | grep basistransform | sed "s/^/hyb$h /" | cut -c1-150 | tee -a gpurun_out/r02r_transform.log
done
for h in 0 2; do
  B2K_TRANSFORM_HYB=$h timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --extra '' > gpurun_out/r02r_bench_hyb$h.json 2> gpurun_out/r02r_bench_hyb$h.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02r_bench_hyb$h.json').read().strip().splitlines()[-1])
    print('hyb$h', round(d['value'],1), 'it/s', {k:(v['avg_ms'],v['frac']) for k,v in d['kernels'].items()}, d['parity']['ok'], d['parity'].get('max_rel_diff_ritz'))
except Exception as e:
    print('hyb$h failed', e); print(open('gpurun_out/r02r_bench_hyb$h.err').read()[-1500:])
PY
done
B2K_COOP_LAUNCH=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --extra '' > gpurun_out/r02r_bench_nocoop.json 2> gpurun_out/r02r_bench_nocoop.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02r_bench_nocoop.json').read().strip().splitlines()[-1])
    print('nocoop', round(d['value'],1), 'it/s', {k:(v['avg_ms'],v['frac']) for k,v in d['kernels'].items()}, d['parity']['ok'])
except Exception as e:
    print('nocoop failed', e); print(open('gpurun_out/r02r_bench_nocoop.err').read()[-1500:])
PY
timeout 600 python -m pytest tests/test_gpu_solvers.py tests/test_gpu_dist.py -m gpu -q --timeout 500 -k "lanczos or chained or dist or sharded" > gpurun_out/r02r_pytest2.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02r_pytest2.log | tail -5
