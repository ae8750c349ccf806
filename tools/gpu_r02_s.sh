#!/bin/bash
# round-2 GPU call S (1 GPU): final state — full GPU suite, full bench line, launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02s_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02s_pytest.log | tail -15
timeout 1200 python bench.py > gpurun_out/r02s_bench_full.json 2> gpurun_out/r02s_bench_full.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02s_bench_full.json').read().strip().splitlines()[-1])
print('FULL', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k:(v['avg_ms'],v['frac']) for k,v in d['kernels'].items()}, d['host_ms_per_step'], d['parity']['ok'], 'launches', d['gpu_launches'])
for k,v in d['other_configs'].items(): print(' ', k, json.dumps(v)[:260])
PY
timeout 600 python tools/run_configs.py cg widened > gpurun_out/r02s_configs.log 2>&1
python - <<PY
import json
d=json.load(open('gpurun_out/configs_cg_widened_n1.json'))
for k,v in d['cg'].items(): print('cg',k,round(v['it_per_s'],1))
w=d['widened']
print('bicgstab', round(w['bicgstab']['ops_per_s'],1), 'blocklanczos', round(w['blocklanczos_p4']['it_per_s'],1), round(w['blocklanczos_p4_fast_block']['it_per_s'],1), 'lsmr', round(w['lsmr']['ops_per_s'],1), 'arnoldi', round(w['arnoldi_eigsolve']['it_per_s'],1), 'expo', round(w['exponentiate']['it_per_s'],1))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02s_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --extra '' > gpurun_out/r02s_ncu_bench.log 2>&1
python tools/launch_shares.py gpurun_out/r02s_launches.csv gpurun_out/r02s_launch_shares.json 2>/dev/null | head -8
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
