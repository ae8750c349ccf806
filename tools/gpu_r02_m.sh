#!/bin/bash
# round-2 GPU call M (1 GPU): stream probe at the library's vector size / slab pitch; block path after the SpMM rewrite,
# the CholeskyQR right-multiply kernel and the batched finalize
mkdir -p gpurun_out
( timeout 60 tools/probes/stream_probe 10000000; timeout 60 tools/probes/stream_probe 10000000 0; timeout 60 tools/probes/stream_probe 10000000 2080; timeout 60 tools/probes/stream_probe 30000000 0 ) 2>&1 | tee gpurun_out/r02m_stream_probe.txt
timeout 600 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_solvers.py -m gpu -q --timeout 300 -k "block or Block or spmm or toric" > gpurun_out/r02m_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02m_pytest.log | tail -8
timeout 120 python tools/block_probe.py fast 3 3 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02m_block.csv python tools/block_probe.py fast 2 1 > gpurun_out/r02m_block.log 2>&1
python tools/launch_shares.py gpurun_out/r02m_block.csv gpurun_out/r02m_block_shares.json | head -12
