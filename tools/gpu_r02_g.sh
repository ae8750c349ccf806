#!/bin/bash
# round-2 GPU call G (1 GPU): full GPU suite on the chosen defaults (chain mode 0, L2 hints), restart-GEMM variants,
# SpMV variant microbench, the full bench line with the other configs, launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02g_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02g_pytest.log | tail -15
for m in 0 1 2 3; do
  B2K_TRANSFORM_UR=$m timeout 200 python tools/microbench.py --reps 8 --k 60 2>&1 | grep basistransform | sed "s/^/ur$m /" | cut -c1-150 | tee -a gpurun_out/r02g_transform.log
done
B2K_SPMV_VARIANT=1 timeout 200 python tools/microbench.py --reps 10 2>&1 | grep spmv | sed "s/^/v1 /" | cut -c1-150 | tee gpurun_out/r02g_spmv_v1.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --extra '' > gpurun_out/r02g_$name.json 2> gpurun_out/r02g_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02g_$name.json').read().strip().splitlines()[-1])
    print('$name', round(d['value'],1), 'it/s', {k:(v['avg_ms'],v['frac']) for k,v in d['kernels'].items()}, d['parity']['ok'])
except Exception as e:
    print('$name failed', e); print(open('gpurun_out/r02g_$name.err').read()[-1500:])
PY
}
run v0 B2K_SPMV_VARIANT=0
run v1 B2K_SPMV_VARIANT=1
run v0_ur1 B2K_TRANSFORM_UR=1
run v0_ur2 B2K_TRANSFORM_UR=2
run v0_ur3 B2K_TRANSFORM_UR=3
timeout 1200 python bench.py > gpurun_out/r02g_bench_full.json 2> gpurun_out/r02g_bench_full.err
tail -c 3000 gpurun_out/r02g_bench_full.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02g_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --extra '' > gpurun_out/r02g_ncu_bench.log 2>&1
tail -3 gpurun_out/r02g_launches.csv | cut -c1-200
