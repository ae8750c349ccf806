import sys, time, numpy as np
sys.path.insert(0, "/root/repo")
import krylovkit_jl_b200 as kk
from oracle import krylov_oracle as ko
SEED = 20260923
for m in (200_000, 2_000_000):
    n = 512
    ctx = kk.B200Context(m, 64, dtype=np.float32)
    sv = ctx.add_space(n, 64, sharded=False)
    op = kk.B200Dense.splitmix(ctx, m, n, SEED, sv)
    u0 = ctx.splitmix(SEED + 1)
    for name, orth, oorth in (("cgs2", kk.cgs2, ko.Orth(ko.CGS2)), ("mgs2", kk.mgs2, ko.Orth(ko.MGS2)),
                              ("cgsr", kk.ClassicalGramSchmidtIR(eta=0.75), ko.Orth(ko.CGSIR, 0.75))):
        alg = kk.GKL(orth=orth, krylovdim=30, maxiter=6, tol=1e-5, verbosity=0)
        t0 = time.time()
        try:
            S, Lv, Rv, info = kk.svdsolve(op, u0, 6, "LR", alg)
            print("GPU", m, name, "S", np.round(S[:6], 3), "numiter", info.numiter, "numops", info.numops, "conv", info.converged,
                  "normres", np.round(info.normres[:3], 4), f"{time.time()-t0:.2f}s", flush=True)
            del Lv, Rv, info
        except Exception as e:
            print("GPU", m, name, "FAILED", type(e).__name__, str(e)[:100], flush=True)
        if m <= 200_000:
            A = ko.dense_splitmix(SEED, m, n)
            uh = ko.splitmix_vector(SEED + 1, m, dtype=np.float32)
            try:
                oS, _, _, oinfo = ko.svdsolve_gkl(A, uh, 6, "LR", krylovdim=30, maxiter=6, tol=1e-5, orth=oorth)
                print("CPU", m, name, "S", np.round(oS[:6], 3), "numiter", oinfo["numiter"], "conv", oinfo["converged"], flush=True)
            except Exception as e:
                print("CPU", m, name, "FAILED", type(e).__name__, str(e)[:100], flush=True)
    if m <= 200_000:
        print("ref svd", np.round(np.linalg.svd(A.astype(np.float64), compute_uv=False)[:6], 3))
    ctx.close()
