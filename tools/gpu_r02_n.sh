#!/bin/bash
# round-2 GPU call N (1 GPU): event-timed breakdown of a CG iteration (no profiler); block path with the vectorised
# CholeskyQR right-multiply
mkdir -p gpurun_out
timeout 200 python tools/cg_breakdown.py 2>&1 | tee gpurun_out/r02n_cg_breakdown.txt
timeout 600 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_solvers.py -m gpu -q --timeout 300 -k "block or Block or spmm or toric" > gpurun_out/r02n_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02n_pytest.log | tail -8
timeout 120 python tools/block_probe.py fast 3 3 2>&1 | tail -2
