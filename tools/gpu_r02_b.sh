#!/bin/bash
# round-2 GPU call B: diagnostics of the chained test (plain, first in process; then under memcheck), full suite
# with the new block / cg-chain / mgs2b tests, the 2-ranks-on-1-GPU sharded test, bench with the extras
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_solvers.py -q -k "chained" > gpurun_out/r02b_chained_plain.log 2>&1
tail -15 gpurun_out/r02b_chained_plain.log
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_solvers.py -q -k "chained" > gpurun_out/r02b_chained_memcheck.log 2>&1
grep -E "assert|Error|passed|failed|=========" gpurun_out/r02b_chained_memcheck.log | tail -15
( time timeout 1200 python -m pytest tests -m gpu -q --maxfail=12 --timeout 300 -k "not one_gpu" ) > gpurun_out/r02b_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02b_pytest.log | tail -30
( time timeout 500 python -m pytest tests/test_gpu_dist.py -x -q -k "one_gpu" ) > gpurun_out/r02b_dist1gpu.log 2>&1
grep -E "Error|assert|passed|failed|dist_check" gpurun_out/r02b_dist1gpu.log | tail -20
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02b_bench.json').read().strip().splitlines()[-1])
    print('bench', round(d['value'],1), 'it/s e2e', round(d['e2e']['value'],1), d['roofline']['frac'], d['parity']['ok'])
    print({k:(v['avg_ms'],v['frac'],v['share_of_step']) for k,v in d['kernels'].items()}, d['gpu_launches'])
    oc=d['other_configs']
    print('c3', {k:(round(v['value'],1)) for k,v in oc['c3'].items() if isinstance(v,dict)}, 'c4', {k:round(v.get('value',0),1) for k,v in oc['c4'].items() if isinstance(v,dict)}, 'c5', round(oc['c5']['value'],1))
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/r02b_bench.err').read()[-2000:])
PY
timeout 300 python bench.py --steps 3 --warmup 3 --orth mgs2b --no-cpu-baseline --no-e2e --extra '' > gpurun_out/r02b_bench_mgs2b.json 2> gpurun_out/r02b_bench_mgs2b.err
timeout 300 python bench.py --steps 3 --warmup 3 --orth mgs2 --no-cpu-baseline --no-e2e --extra '' > gpurun_out/r02b_bench_mgs2.json 2> gpurun_out/r02b_bench_mgs2.err
python - <<'PY'
import json
for t in ('mgs2b','mgs2'):
    try:
        d=json.loads(open(f'gpurun_out/r02b_bench_{t}.json').read().strip().splitlines()[-1]); print(t, round(d['value'],1), d['ritz_values'][:2], d.get('parity'))
    except Exception as e:
        print(t,'failed',e); print(open(f'gpurun_out/r02b_bench_{t}.err').read()[-1500:])
PY
