#!/bin/bash
# round-2 GPU call C (N GPUs of one box; N from $1, default 2): sharded parity tests, then the strong-scaling
# bench at N with the NVLink peer window (default) and with NCCL only (B2K_PEER=0) for comparison
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
if [ -z "$SKIP_TESTS" ]; then
bash tools/gpu_r02_e.sh
( time timeout 600 python -m pytest tests/test_gpu_dist.py -q --timeout 500 ) > gpurun_out/r02c_dist_n$N.log 2>&1
grep -E "Error|assert|passed|failed|dist_check" gpurun_out/r02c_dist_n$N.log | tail -12
fi
P=29800
for MODE in peer nccl; do
  P=$((P+1))
  if [ $MODE = nccl ]; then export B2K_PEER=0; else unset B2K_PEER; fi
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 5 --warmup 3 --extra "$EXTRA" > gpurun_out/r02c_bench_n${N}_$MODE.json 2> gpurun_out/r02c_bench_n${N}_$MODE.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02c_bench_n${N}_$MODE.json').read().strip().splitlines()[-1])
    oc=d.get('other_configs') or {}
    if oc: print('  other:', {k:(round(v['value'],1) if 'value' in v else {kk_:round(vv['value'],1) for kk_,vv in v.items() if isinstance(vv,dict) and 'value' in vv}) for k,v in oc.items()})
    k=d['kernels']; print('  kernel ms/job', round(sum(v['ms_total'] for v in k.values())/d['steps'],2), 'launches', d['gpu_launches'])
    print('$MODE N=$N', round(d['value'],1), 'it/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1), d['ritz_values'][:2], d['parity']['ok'], {k:(v['avg_ms'],v['frac']) for k,v in d['kernels'].items()})
except Exception as e:
    print('$MODE failed', e); print(open('gpurun_out/r02c_bench_n${N}_$MODE.err').read()[-2500:])
PY
done
unset B2K_PEER
