#!/bin/bash
# round-2 GPU call F (1 GPU): A/B of the SpMV variants and the L2 hints on the headline job, host section profile
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --extra '' > gpurun_out/r02f_$name.json 2> gpurun_out/r02f_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02f_$name.json').read().strip().splitlines()[-1])
    print('$name', round(d['value'],1), 'it/s', {k:(v['avg_ms'],v['frac']) for k,v in d['kernels'].items()}, d.get('host_ms_per_step'), d['parity']['ok'])
except Exception as e:
    print('$name failed', e); print(open('gpurun_out/r02f_$name.err').read()[-1500:])
PY
}
run m1_h1_v0 B2K_CHAIN_MODE=1 B2K_L2_HINTS=1 B2K_SPMV_VARIANT=0
run m1_h0_v0 B2K_CHAIN_MODE=1 B2K_L2_HINTS=0 B2K_SPMV_VARIANT=0
run m0_h1_v0 B2K_CHAIN_MODE=0 B2K_L2_HINTS=1 B2K_SPMV_VARIANT=0
run m0_h0_v0 B2K_CHAIN_MODE=0 B2K_L2_HINTS=0 B2K_SPMV_VARIANT=0
run m1_h1_v1 B2K_CHAIN_MODE=1 B2K_L2_HINTS=1 B2K_SPMV_VARIANT=1
run m0_h1_v1 B2K_CHAIN_MODE=0 B2K_L2_HINTS=1 B2K_SPMV_VARIANT=1
run m1_h1_v0_again B2K_CHAIN_MODE=1 B2K_L2_HINTS=1 B2K_SPMV_VARIANT=0
timeout 200 python tools/microbench.py --reps 10 > gpurun_out/r02f_microbench.log 2>&1
cut -c1-160 gpurun_out/r02f_microbench.log
timeout 300 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_solvers.py -q --timeout 200 -k "block or chained or cg_chained or blocklanczos or matrix_free" > gpurun_out/r02f_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02f_pytest.log | tail
