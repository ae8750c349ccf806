#!/bin/bash
# FIRST GPU call for the one-pass GKL step (written after round 2's GPU budget had ended; never run on a B200):
#   1. its GPU tests (both kernel variants, the solver in both precisions),
#   2. the bench record that times the two-pass step against the one-pass step and both kernel variants (child mode),
#   3. an ncu --set full capture of each fused kernel (DRAM throughput with 128- / 256-byte column segments, stall reasons),
#   4. compute-sanitizer memcheck + racecheck on a small shape (plain SIMT kernels: racecheck models them fully).
# Usage:  /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_next_onepass.sh'
mkdir -p gpurun_out
t0=$(date +%s)
timeout 300 python -m pytest tests/test_gpu_zzz_onepass.py tests/test_gpu_fullsize.py -q -m gpu -k "onepass or gram or config4" > gpurun_out/onepass_tests.log 2>&1
echo "tests rc=$? $(( $(date +%s) - t0 )) s: $(tail -1 gpurun_out/onepass_tests.log)"
B2K_BENCH_CHILD=1 timeout 300 python bench.py --extra c4o > gpurun_out/onepass_bench.json 2> gpurun_out/onepass_bench.err
echo "bench rc=$? $(( $(date +%s) - t0 )) s"; cut -c1-1500 gpurun_out/onepass_bench.json
cat > /tmp/onepass_once.py <<'PY'
import sys, numpy as np
import krylovkit_jl_b200 as kk
v = int(sys.argv[1]); m = int(sys.argv[2])
lib = kk._lib.load(); lib.b2k_debug_set_onepass_variant(v)
ctx = kk.B200Context(m, 6, dtype=np.float32); sv = ctx.add_space(512, 8, sharded=False)
op = kk.B200Dense.splitmix(ctx, m, 512, 20260923, sv); x = ctx.splitmix(7, sv)
for _ in range(3):
    y, z = kk.apply_normal_gram(op, x); del y, z
lib.b2k_device_sync(); print("ok")
PY
for v in 0 1; do
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_dense_onepass -s 2 -c 1 -o gpurun_out/onepass_v$v python /tmp/onepass_once.py $v 2000000 > gpurun_out/onepass_ncu_v$v.log 2>&1
  echo "ncu v$v rc=$? $(( $(date +%s) - t0 )) s"
done
for tool in memcheck racecheck; do
  for v in 0 1; do
    timeout 200 compute-sanitizer --tool $tool python /tmp/onepass_once.py $v 4000 > gpurun_out/onepass_${tool}_v$v.log 2>&1
    echo "$tool v$v rc=$?: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/onepass_${tool}_v$v.log | tail -1)"
  done
done
echo "total $(( $(date +%s) - t0 )) s"
