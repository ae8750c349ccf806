#!/bin/bash
# round-2 GPU call U (1 GPU): ncu --set full capture of the DFMA 8x9 restart GEMM (pipe utilisation, stall reasons)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_transform_f89 -s 2 -c 1 -o gpurun_out/r02u_f89 -f python tools/microbench.py --reps 4 --k 60 > gpurun_out/r02u_ncu.log 2>&1
tail -3 gpurun_out/r02u_ncu.log
ncu -i gpurun_out/r02u_f89.ncu-rep --page raw --csv 2>/dev/null | python - <<'PY'
import csv, sys
rows=list(csv.reader(sys.stdin))
if len(rows) < 3: print("no rows"); sys.exit()
h=rows[0]; r=rows[-1]
want=["gpu__time_duration.sum","sm__inst_executed_pipe_fp64","sm__pipe_fp64_cycles_active","smsp__inst_executed_pipe_fp64","l1tex__data_pipe_lsu_wavefronts_mem_shared","l1tex__data_bank_conflicts","dram__throughput","dram__bytes_read.sum","dram__bytes_write.sum","sm__warps_active","smsp__issue_active","smsp__average_warp","sm__throughput","launch__registers_per_thread","smsp__cycles_active.avg","sm__cycles_elapsed.max","smsp__inst_executed.sum"]
for i,name in enumerate(h):
    if any(w in name for w in want): print(name, '=', r[i])
PY
python tools/ncu_top.py gpurun_out/r02u_f89.ncu-rep 12 2>/dev/null | head -20
