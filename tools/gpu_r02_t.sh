#!/bin/bash
# round-2 GPU call T (repeat of H after the partial-sum change) (8 GPUs of one box): the strong-scaling table of the headline job on ONE box — N = 8, 4, 2, 1
# with the NVLink peer window, N = 8 also with NCCL only (B2K_PEER=0) — and config 5 (512^3 7-point) at N = 8
mkdir -p gpurun_out
nvidia-smi -L | head -8
P=29900
run() {  # N, mode, extra
  N=$1; MODE=$2; EX=$3
  P=$((P+1))
  if [ $MODE = nccl ]; then export B2K_PEER=0; else unset B2K_PEER; fi
  out=gpurun_out/r02t_bench_n${N}_$MODE
  if [ $N = 1 ]; then
    timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --extra "$EX" > $out.json 2> $out.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 5 --warmup 3 --no-e2e --extra "$EX" > $out.json 2> $out.err
  fi
  python - <<PY
import json
try:
    d=json.loads(open('$out.json').read().strip().splitlines()[-1])
    oc=d.get('other_configs') or {}
    if oc: print('  other:', json.dumps(oc)[:600])
    k=d['kernels']; print('  kernel ms/job', round(sum(v['ms_total'] for v in k.values())/d['steps'],2), 'launches', d['gpu_launches'])
    print('$MODE N=$N', round(d['value'],1), 'it/s', round(d['ms_per_step'],2), 'ms', d['ritz_values'][:2], d['parity']['ok'], {k:(v['avg_ms'],v['frac']) for k,v in d['kernels'].items()})
except Exception as e:
    print('$MODE N=$N failed', e); print(open('$out.err').read()[-2500:])
PY
  unset B2K_PEER
}
run 8 peer c5
run 4 peer ''
run 2 peer ''
run 1 peer ''
