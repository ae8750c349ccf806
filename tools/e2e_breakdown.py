"""Wall-clock breakdown of the host-buffer eigsolve path (context create / uploads / solve / downloads)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import krylovkit_jl_b200 as kk  # noqa: E402
from bench import pinned_array  # noqa: E402

nx, ny, kd = 4000, 2500, 60
n = nx * ny
lib = kk._lib.load()
ctx0 = kk.B200Context(n, 8)
op0 = kk.B200CSR.stencil(ctx0, nx, ny)
rp, _ = pinned_array(lib, n + 1, np.int32)
ci, _ = pinned_array(lib, op0.nnz, np.int32)
va, _ = pinned_array(lib, op0.nnz, np.float64)
ctx0.check(lib.b2k_op_csr_download(ctx0.h, op0.h, rp.ctypes.data, ci.ctypes.data, va.ctypes.data))
xh, _ = pinned_array(lib, n, np.float64)
ctx0.splitmix(1).to_host(xh)
outs = [pinned_array(lib, n, np.float64)[0] for _ in range(4)]
ctx0.close()


def T():
    lib.b2k_device_sync()
    return time.perf_counter()


for rep in range(3):
    t = [T()]
    ctx = kk.B200Context(n, kd + 16); t.append(T())
    op = kk.B200CSR.from_csr_arrays(ctx, n, n, rp, ci, va); t.append(T())
    x0 = ctx.from_host(xh); t.append(T())
    alg = kk.Lanczos(orth=kk.cgs2, krylovdim=kd, maxiter=5, tol=0.0, verbosity=0)
    vals, vecs, info = kk.eigsolve(op, x0, 4, "SR", alg); t.append(T())
    for i, v in enumerate(vecs):
        v.to_host(outs[i])
    t.append(T())
    del vecs, info, x0, op
    ctx.close(); t.append(T())
    names = ["ctx_create", "op_create(H2D 640MB)", "x0 upload", "solve", "download 4 vecs", "ctx close"]
    print(rep, {k: round(1000 * (b - a), 1) for k, a, b in zip(names, t[:-1], t[1:])}, "total", round(1000 * (t[-1] - t[0]), 1))

# the exact bench.py e2e call, with a second (resident) context alive like in bench.py
ctx_main = kk.B200Context(n, kd + 16)
op_main = kk.B200CSR.stencil(ctx_main, nx, ny)
alg = kk.Lanczos(orth=kk.cgs2, krylovdim=kd, maxiter=5, tol=0.0, verbosity=0)
import cProfile, pstats
for rep in range(4):
    t0 = T()
    if rep == 3:
        pr = cProfile.Profile(); pr.enable()
    vals, vecs, info = kk.eigsolve((rp, ci, va), xh, 4, "SR", alg, out_vectors=outs)
    if rep == 3:
        pr.disable()
    t1 = T()
    print("host-buffer eigsolve", rep, round(1000 * (t1 - t0), 1), "ms")
pstats.Stats(pr).sort_stats("tottime").print_stats(12)
