#!/bin/bash
# round-2 GPU call I (1 GPU): SpMV variant 1 as default under the kernel/solver tests, the chained BiCGStab test,
# and the widened rows at the 1e7 scale (CG / BiCGStab chains, BlockLanczos reference vs fast block mode)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_solvers.py tests/test_cclient.py -m gpu -q --timeout 600 > gpurun_out/r02i_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02i_pytest.log | tail -15
timeout 600 python tools/run_configs.py cg widened > gpurun_out/r02i_configs.log 2>&1
tail -c 6000 gpurun_out/r02i_configs.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --extra '' > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02i_bench.json').read().strip().splitlines()[-1])
print(round(d['value'],1), 'it/s', {k:(v['avg_ms'],v['frac']) for k,v in d['kernels'].items()}, d.get('host_ms_per_step'), d['parity']['ok'])
PY
