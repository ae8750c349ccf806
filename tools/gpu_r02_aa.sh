#!/bin/bash
# round-2 GPU call AA (1 GPU, ~2.5 min of budget left): the peer-wait watchdog — the sharded path still passes with
# peer_spin in the kernels (two ranks on one GPU, short form), the watchdog test itself, then smoke()
mkdir -p gpurun_out
t0=$(date +%s)
B2K_ONE_GPU=1 DIST_CHECK_SHORT=1 timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tools/dist_check.py > gpurun_out/r02aa_short.log 2>&1
echo "short rc=$? $(( $(date +%s) - t0 )) s: $(grep 'dist_check ok' gpurun_out/r02aa_short.log | cut -c1-120)"
B2K_ONE_GPU=1 DIST_CHECK_WATCHDOG=1 timeout 40 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 tools/dist_check.py > gpurun_out/r02aa_watchdog.log 2>&1
echo "watchdog rc=$? $(( $(date +%s) - t0 )) s: $(grep 'dist_check ok' gpurun_out/r02aa_watchdog.log | cut -c1-120)"
tail -5 gpurun_out/r02aa_watchdog.log | cut -c1-300
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "total $(( $(date +%s) - t0 )) s"
