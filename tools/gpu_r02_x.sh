#!/bin/bash
# round-2 GPU call X (1 GPU): restart-GEMM tests after the posted-use guard, event-trace test, one-GPU step timeline
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_primitives.py -m gpu -q --timeout 120 -x -k "basistransform" > gpurun_out/r02x_pytest.log 2>&1
tail -3 gpurun_out/r02x_pytest.log | cut -c1-200
timeout 200 python -m pytest tests/test_gpu_solvers.py -m gpu -q --timeout 120 -k "event_trace or chained" 2>&1 | tail -3
B2K_TRANSFORM_HYB=2 timeout 100 python tools/microbench.py --reps 8 --k 60 2>&1 | grep basistransform | cut -c1-150
timeout 200 python tools/trace_step.py 2>&1 | tail -16
