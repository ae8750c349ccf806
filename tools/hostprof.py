"""cProfile of one benchmark job (host-side overhead between kernel launches)."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import krylovkit_jl_b200 as kk  # noqa: E402

nx, ny, kd = 4000, 2500, 60
n = nx * ny
ctx = kk.B200Context(n, kd + 16)
op = kk.B200CSR.stencil(ctx, nx, ny)
x0 = ctx.splitmix(20260923)
alg = kk.Lanczos(orth=kk.cgs2, krylovdim=kd, maxiter=5, tol=0.0, verbosity=0)
kk.eigsolve(op, x0, 4, "SR", alg)
t0 = time.perf_counter()
kk.eigsolve(op, x0, 4, "SR", alg)
print("job wall", time.perf_counter() - t0)
pr = cProfile.Profile()
pr.enable()
kk.eigsolve(op, x0, 4, "SR", alg)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(18)
