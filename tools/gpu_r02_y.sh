#!/bin/bash
# round-2 GPU call Y (1 GPU): fence + relaxed flag stores under the sharded tests (2 ranks on one GPU), restart GEMM
# variants (incl. the 512-row-tile one), event trace, quick bench
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 400 -k "two_ranks_on_one_gpu" > gpurun_out/r02y_dist.log 2>&1
tail -3 gpurun_out/r02y_dist.log | cut -c1-200
timeout 300 python -m pytest tests/test_gpu_primitives.py -m gpu -q --timeout 120 -x -k "basistransform" 2>&1 | tail -2
timeout 200 python -m pytest tests/test_gpu_solvers.py -m gpu -q --timeout 120 -k "event_trace or chained" 2>&1 | tail -2
for h in 2 3; do
  B2K_TRANSFORM_HYB=$h timeout 100 python tools/microbench.py --reps 8 --k 60 2>&1 | grep basistransform | sed "s/^/hyb$h /" | cut -c1-150 | tee -a gpurun_out/r02y_transform.log
done
for h in 2 3; do
B2K_TRANSFORM_HYB=$h timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --extra '' > gpurun_out/r02y_bench_hyb$h.json 2> gpurun_out/r02y_bench_hyb$h.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02y_bench_hyb$h.json').read().strip().splitlines()[-1])
    print('hyb$h', round(d['value'],1), 'it/s', {k:(v['avg_ms'],v['frac']) for k,v in d['kernels'].items()}, d['parity']['ok'])
except Exception as e:
    print('hyb$h failed', e); print(open('gpurun_out/r02y_bench_hyb$h.err').read()[-1500:])
PY
done
