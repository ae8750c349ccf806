"""BlockLanczos (p = 4, krylovdim 32) on the 1e7 5-point Laplacian, fast block mode or reference mode: meant to run under
`ncu --metrics gpu__time_duration.sum` for a per-launch list, or plain for the host-section profile."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import krylovkit_jl_b200 as kk  # noqa: E402

fast = (sys.argv[1] if len(sys.argv) > 1 else "fast") == "fast"
cycles = int(sys.argv[2]) if len(sys.argv) > 2 else 2
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
nx, ny = 4000, 2500
ctx = kk.B200Context(nx * ny, 64)
op = kk.B200CSR.stencil(ctx, nx, ny)
X0 = kk.Block([ctx.splitmix(7 + i) for i in range(4)])
alg = kk.BlockLanczos(krylovdim=32, maxiter=cycles, tol=0.0, verbosity=0, fast_block=fast)
for r in range(reps):
    ctx.lib.b2k_device_sync()
    t0 = time.perf_counter()
    vals, vecs, info = kk.eigsolve(op, X0, 4, "SR", alg)
    ctx.lib.b2k_device_sync()
    dt = time.perf_counter() - t0
    print("fast" if fast else "reference", info.numops, info.numiter, [float(v) for v in vals[:4]], f"{dt * 1e3:.1f} ms wall",
          f"{info.numops / dt:.1f} it/s")
    del vecs
ctx.close()
