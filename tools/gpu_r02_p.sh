#!/bin/bash
# round-2 GPU call P (1 GPU): everything again WITHOUT the persisting-L2 set-aside (the new default): full GPU suite,
# full bench line, literal MGS2 with hints vs with the old set-aside + window, widened solvers, launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02p_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02p_pytest.log | tail -15
timeout 600 python tools/run_configs.py cg widened > gpurun_out/r02p_configs.log 2>&1
python - <<PY
import json
d=json.load(open('gpurun_out/configs_cg_widened_n1.json'))
for k,v in d['cg'].items(): print('cg',k,round(v['it_per_s'],1))
w=d['widened']
print('bicgstab', round(w['bicgstab']['ops_per_s'],1), 'blocklanczos', round(w['blocklanczos_p4']['it_per_s'],1), round(w['blocklanczos_p4_fast_block']['it_per_s'],1), 'lsmr', round(w['lsmr']['ops_per_s'],1), 'arnoldi', round(w['arnoldi_eigsolve']['it_per_s'],1), 'expo', round(w['exponentiate']['it_per_s'],1))
PY
run() {  # name, args..., env via prefix
  name=$1; shift
  timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --extra '' "$@" > gpurun_out/r02p_$name.json 2> gpurun_out/r02p_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02p_$name.json').read().strip().splitlines()[-1])
    print('$name', round(d['value'],1), 'it/s', {k:(v['avg_ms'],v['frac']) for k,v in d['kernels'].items()}, d['parity']['ok'])
except Exception as e:
    print('$name failed', e); print(open('gpurun_out/r02p_$name.err').read()[-1500:])
PY
}
run mgs2_hints --orth mgs2
B2K_L2_CARVE=1000000000 run mgs2_carve --orth mgs2
run mgs2b --orth mgs2b
timeout 1200 python bench.py > gpurun_out/r02p_bench_full.json 2> gpurun_out/r02p_bench_full.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02p_bench_full.json').read().strip().splitlines()[-1])
print('FULL', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k:(v['avg_ms'],v['frac']) for k,v in d['kernels'].items()}, d['host_ms_per_step'], d['parity']['ok'])
for k,v in d['other_configs'].items(): print(' ', k, json.dumps(v)[:300])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02p_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --extra '' > gpurun_out/r02p_ncu_bench.log 2>&1
python tools/launch_shares.py gpurun_out/r02p_launches.csv gpurun_out/r02p_launch_shares.json | head -8
