#!/bin/bash
# round-2 GPU call A: full GPU test-suite, chained-Lanczos sanitizer pass, short bench, BLAS-1 microbench,
# then the sharded path with two ranks on the one GPU
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
( time timeout 1200 python -m pytest tests -m gpu -q --maxfail=12 --timeout 300 -k "not one_gpu" ) > gpurun_out/r02a_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02a_pytest.log | tail -30
timeout 300 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_solvers.py -x -q -k "chained" > gpurun_out/r02a_racecheck.log 2>&1
tail -8 gpurun_out/r02a_racecheck.log
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_solvers.py -x -q -k "chained or invariant" > gpurun_out/r02a_memcheck.log 2>&1
tail -5 gpurun_out/r02a_memcheck.log
timeout 400 python bench.py --steps 5 --warmup 3 --extra '' > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02a_bench.json').read().strip().splitlines()[-1])
    print('bench', round(d['value'],1), 'it/s e2e', round(d['e2e']['value'],1), d['ritz_values'], d['roofline']['frac'], d.get('parity'))
    print({k:(v['avg_ms'],v['frac'],v['share_of_step']) for k,v in d['kernels'].items()}, d['gpu_launches'])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/r02a_bench.err').read()[-2000:])
PY
timeout 200 python tools/microbench.py --reps 10 > gpurun_out/r02a_microbench.log 2>&1
grep -E "scale|axpby|nrm2|inner|dot|axpy" gpurun_out/r02a_microbench.log | cut -c1-200
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02a_bench_extras.json 2> gpurun_out/r02a_bench_extras.err
tail -c 2500 gpurun_out/r02a_bench_extras.json; tail -5 gpurun_out/r02a_bench_extras.err
( time timeout 500 python -m pytest tests/test_gpu_dist.py -x -q -k "one_gpu" ) > gpurun_out/r02a_dist1gpu.log 2>&1
tail -30 gpurun_out/r02a_dist1gpu.log
