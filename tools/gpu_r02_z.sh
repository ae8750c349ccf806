#!/bin/bash
# round-2 GPU call Z (8 GPUs): N = 8 after the relaxed-flag publishes — bench line and step timeline
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29791 bench.py --gpus 8 --steps 5 --warmup 3 --no-e2e --extra '' > gpurun_out/r02z_bench_n8.json 2> gpurun_out/r02z_bench_n8.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02z_bench_n8.json').read().strip().splitlines()[-1])
    print('N=8', round(d['value'],1), 'it/s', round(d['ms_per_step'],2), 'ms', d['parity']['ok'], {k:(v['avg_ms'],v['frac']) for k,v in d['kernels'].items()})
except Exception as e:
    print('failed', e); print(open('gpurun_out/r02z_bench_n8.err').read()[-2000:])
PY
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29792 tools/trace_step.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$\|NCCL version\|destroy_process" | head -14
