#!/bin/bash
# round-2 GPU call D (1 GPU): the three chain tests (plain + memcheck + racecheck of the cooperative kernels),
# the BLAS-1 streaming probe, an ncu launch list of one job and --set full captures of the chained SpMV and GS kernel
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_solvers.py -q -k "chained" > gpurun_out/r02d_chained_plain.log 2>&1
grep -E "assert|Error|passed|failed" gpurun_out/r02d_chained_plain.log | tail -8
timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_solvers.py -q -k "chained" > gpurun_out/r02d_chained_memcheck.log 2>&1
grep -E "assert|Error|passed|failed|=========" gpurun_out/r02d_chained_memcheck.log | tail -8
timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_primitives.py -q -k "orthogonalize or basistransform or block_multi" > gpurun_out/r02d_racecheck_gs_transform_block.log 2>&1
grep -E "RACECHECK|passed|failed|hazard" gpurun_out/r02d_racecheck_gs_transform_block.log | tail -8
./tools/probes/blas1_probe > gpurun_out/r02d_blas1_probe.log 2>&1; cat gpurun_out/r02d_blas1_probe.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 400 --csv --log-file gpurun_out/r02d_launches.csv python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-e2e --extra '' > gpurun_out/r02d_ncu_bench.log 2>&1
python tools/launch_shares.py gpurun_out/r02d_launches.csv gpurun_out/r02d_launch_shares.json | head -20
ncu --set full --clock-control none --import-source on -k regex:k_spmv_pipe -s 110 -c 1 -o gpurun_out/r02d_prof_spmv -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --extra '' > gpurun_out/r02d_ncu_spmv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_gs_fused -s 100 -c 1 -o gpurun_out/r02d_prof_gs -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --extra '' > gpurun_out/r02d_ncu_gs.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
