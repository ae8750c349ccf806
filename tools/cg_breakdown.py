"""Where a CG iteration's time goes at n = 1e7 WITHOUT a profiler: CUDA-event timings (library timer) of each kernel of
the iteration repeated on its own, of the pairs, and of the chained iteration."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import krylovkit_jl_b200 as kk  # noqa: E402

nx, ny = 4000, 2500
n = nx * ny
ctx = kk.B200Context(n, 16)
lib = ctx.lib
op = kk.B200CSR.stencil(ctx, nx, ny)
x, r, p, q = ctx.full(0.0), ctx.splitmix(1), ctx.splitmix(2), ctx.full(0.0)
W = 8.0 * n
spmv = op.nnz * 12 + 4 * (n + 1) + 2 * W


def timed(name, fn, reps, nbytes):
    for _ in range(5):
        fn()
    lib.b2k_device_sync()
    lib.b2k_timer_start(ctx.h)
    for _ in range(reps):
        fn()
    ms = C.c_double()
    lib.b2k_timer_stop(ctx.h, C.byref(ms))
    t = ms.value / reps
    print(f"{name:44s} {t * 1e3:8.1f} us  {nbytes / t / 1e6:7.1f} GB/s", flush=True)
    return t


d = C.c_double()
pq, nr = C.c_double(), C.c_double()
axpby = lambda: lib.b2k_vec_axpby(ctx.h, p.handle, r.handle, 1.0, 0.5)           # p <- r + 0.5 p   (2r + 1w)
apply_dot = lambda: lib.b2k_op_apply_dot(ctx.h, op.h, p.handle, q.handle, p.handle, C.byref(d))
apply = lambda: lib.b2k_op_apply(ctx.h, op.h, p.handle, q.handle)
cg_step = lambda: lib.b2k_cg_step(ctx.h, op.h, x.handle, r.handle, p.handle, q.handle, 0.0, 1.0, 0.5, 1.0, C.byref(pq), C.byref(nr))
timed("axpby p <- r + b p (3W)", axpby, 200, 3 * W)
timed("apply (SpMV)", apply, 200, spmv)
timed("apply_dot (SpMV + <p,q>, one host sync)", apply_dot, 200, spmv)
timed("axpby ; apply", lambda: (axpby(), apply()), 200, 3 * W + spmv)
timed("cg_step (axpby ; SpMV+dot ; x,r update; sync)", cg_step, 200, 9 * W + spmv)
pqs, nrs, done = (C.c_double * 64)(), (C.c_double * 64)(), C.c_int32()
chain = lambda: lib.b2k_cg_chain(ctx.h, op.h, x.handle, r.handle, p.handle, q.handle, 0.0, 1.0, 0.5, 1.0, 0.0, 32, pqs, nrs, C.byref(done))
t = timed("cg_chain, 32 iterations per call", chain, 10, 32 * (9 * W + spmv))
print(f"  -> {32 / t:.0f} it/s")
ctx.close()
