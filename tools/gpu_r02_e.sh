#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_solvers.py -q -k "chained" > gpurun_out/r02e_chained_plain.log 2>&1
grep -E "assert|Error|passed|failed" gpurun_out/r02e_chained_plain.log | tail -8
timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_solvers.py -q -x -s -k "chained" > gpurun_out/r02e_chained_memcheck.log 2>&1
grep -E "AssertionError|b200krylov|passed|failed|=========" gpurun_out/r02e_chained_memcheck.log | tail -12
