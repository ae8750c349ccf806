#!/bin/bash
# round-2 GPU call K (1 GPU): load-batched BLAS-1 kernels + two-row matrix-free stencil under the kernel/solver tests,
# CG / BiCGStab / BlockLanczos rates again, launch list of a fast-block BlockLanczos run
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_solvers.py tests/test_cclient.py -m gpu -q --timeout 600 > gpurun_out/r02k_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02k_pytest.log | tail -15
timeout 600 python tools/run_configs.py cg widened > gpurun_out/r02k_configs.log 2>&1
python - <<PY
import json
d=json.load(open('gpurun_out/configs_cg_widened_n1.json'))
for k,v in d['cg'].items(): print('cg',k,round(v['it_per_s'],1))
w=d['widened']
print('bicgstab', round(w['bicgstab']['ops_per_s'],1), 'blocklanczos', round(w['blocklanczos_p4']['it_per_s'],1), round(w['blocklanczos_p4_fast_block']['it_per_s'],1), 'lsmr', round(w['lsmr']['ops_per_s'],1), 'arnoldi', round(w['arnoldi_eigsolve']['it_per_s'],1), 'expo', round(w['exponentiate']['it_per_s'],1))
PY
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --extra 'c2f' > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02k_bench.json').read().strip().splitlines()[-1])
print(round(d['value'],1), 'it/s', {k:(v['avg_ms'],v['frac']) for k,v in d['kernels'].items()}, d.get('host_ms_per_step'), d['parity']['ok'], 'c2f', round(d['other_configs']['c2_matrix_free']['value'],1))
PY
timeout 120 python tools/block_probe.py fast 3 3 2>&1 | tail -3
timeout 120 python tools/block_probe.py reference 3 2 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02k_block.csv python tools/block_probe.py fast 2 1 > gpurun_out/r02k_block.log 2>&1
python tools/launch_shares.py gpurun_out/r02k_block.csv gpurun_out/r02k_block_shares.json | head -25
