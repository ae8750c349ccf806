#!/bin/bash
# round-2 GPU call V (1 GPU): two-set DFMA 8x9 restart GEMM — tests, standalone time, job, ncu pipe counters
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_primitives.py -m gpu -q --timeout 300 -k "basistransform" > gpurun_out/r02v_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02v_pytest.log | tail -8
for h in 0 2; do
  B2K_TRANSFORM_HYB=$h timeout 200 python tools/microbench.py --reps 8 --k 60 2>&1 | grep basistransform | sed "s/^/hyb$h /" | cut -c1-150 | tee -a gpurun_out/r02v_transform.log
done
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --extra '' > gpurun_out/r02v_bench.json 2> gpurun_out/r02v_bench.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02v_bench.json').read().strip().splitlines()[-1])
    print('bench', round(d['value'],1), 'it/s', {k:(v['avg_ms'],v['frac']) for k,v in d['kernels'].items()}, d['parity']['ok'], d['parity'].get('max_rel_diff_ritz'))
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/r02v_bench.err').read()[-1500:])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_transform_f89 -s 2 -c 1 -o gpurun_out/r02v_f89 -f python tools/microbench.py --reps 4 --k 60 > gpurun_out/r02v_ncu.log 2>&1
tail -2 gpurun_out/r02v_ncu.log
timeout 900 python -m pytest tests/test_gpu_solvers.py tests/test_gpu_fullsize.py -m gpu -q --timeout 600 -k "eigsolve or lanczos or fullsize" > gpurun_out/r02v_pytest2.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02v_pytest2.log | tail -5
timeout 200 python tools/trace_step.py 2>&1 | tail -18
timeout 300 python -m pytest tests/test_gpu_solvers.py -m gpu -q --timeout 200 -k "event_trace" 2>&1 | tail -3
