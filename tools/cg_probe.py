"""CG / BiCGStab on the 1e7 5-point problem, few iterations: meant to run under
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` for a per-launch list of the
short-recurrence solvers' kernels in their real context (nothing L2-resident between launches)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import krylovkit_jl_b200 as kk  # noqa: E402

nx, ny = 4000, 2500
ctx = kk.B200Context(nx * ny, 16)
op = kk.B200CSR.stencil(ctx, nx, ny)
b = kk.apply(op, ctx.full(1.0))
which = sys.argv[1] if len(sys.argv) > 1 else "cg"
its = int(sys.argv[2]) if len(sys.argv) > 2 else 24
if which == "cg":
    x, info = kk.linsolve(op, b, None, kk.CG(maxiter=its, tol=1e-300, verbosity=0))
else:
    cd5 = kk.B200CSR.stencil(ctx, nx, ny, 1, (5.0, -1.4, -0.6, -1.2, -0.8, 0.0, 0.0))
    x, info = kk.linsolve(cd5, ctx.splitmix(3), None, kk.BiCGStab(maxiter=its, tol=1e-300, verbosity=0))
print(which, info.numiter, info.numops, float(info.normres))
ctx.close()
