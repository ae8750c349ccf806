#!/bin/bash
# strong-scaling sweep of bench.py (config 2, 1e7 rows) and config 5 (8e7 rows) on one 8-GPU box
mkdir -p gpurun_out
P=29700
for N in 1 2 4 8; do
  P=$((P+1))
  if [ $N -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>gpurun_out/scale_err_$N.log | tail -1 > gpurun_out/scale_bench_$N.json
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 3 --warmup 2 2>gpurun_out/scale_err_$N.log | tail -1 > gpurun_out/scale_bench_$N.json
  fi
  python -c "
import json
d=json.load(open('gpurun_out/scale_bench_$N.json')); print('bench N=$N', round(d['value'],1), 'it/s', round(d['ms_per_step'],1), 'ms', d['ritz_values'][:2])" || tail -5 gpurun_out/scale_err_$N.log
done
for N in 8 4 2; do
  P=$((P+1))
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P tools/run_configs.py c5 2>gpurun_out/scale_c5_err_$N.log | grep "^c5" | cut -c1-400 || tail -5 gpurun_out/scale_c5_err_$N.log
done
