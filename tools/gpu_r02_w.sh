#!/bin/bash
# round-2 GPU call W (8 GPUs): event timeline of the chained step at N = 8 (and N = 4 on the same box)
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29781 tools/trace_step.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -40
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29782 tools/trace_step.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -20
