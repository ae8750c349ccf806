"""The product's own svdsolve driver on the numpy stand-in (tests/hostsim.py) at config 4's FULL size — dense 2e6 x 512 Float32
splitmix matrix, GKL krylovdim 30, tol 1e-5, 6 triplets :LR — in the reference's two-pass form and in the flagged one-pass form
(GKL(onepass=True)), MGS2 and CGSIR: numops, restart cycles, passes over A, converged sigma against the Float64 truth.
CPU only, ~15 min and ~6 GB; analysis / test infrastructure, not the product path.  Output of the run behind DESIGN.md §6:
profiles/r02_onepass_fullsize_sim.txt.

    python tools/onepass_fullsize_sim.py 2000000
"""
import sys, time; import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import hostsim
import krylovkit_jl_b200 as kk
from oracle import krylov_oracle as ko
SEED = 20260923
m, n = int(sys.argv[1]), 512
t0=time.time()
A = ko.dense_splitmix(SEED, m, n)
print("gen", time.time()-t0, flush=True)
G64 = np.zeros((n, n))
for i in range(0, m, 100_000):
    blk = A[i:i + 100_000].astype(np.float64); G64 += blk.T @ blk
truth = np.sqrt(np.sort(np.linalg.eigvalsh(G64))[::-1][:6])
print("truth", truth, flush=True)
with hostsim.installed():
    ctx = kk.B200Context(m, 3*30+12, dtype=np.float32)
    sv = ctx.add_space(n, 110, sharded=False)
    op = kk.B200Dense.splitmix(ctx, m, n, SEED, sv)
    u0 = ctx.splitmix(SEED + 1)
    for orth, name in ((kk.mgs2, "mgs2"), (kk.ClassicalGramSchmidtIR(eta=0.75), "cgsr")):
        for onepass in (False, True):
            t1=time.time()
            alg = kk.GKL(orth=orth, krylovdim=30, maxiter=100, tol=1e-5, verbosity=0, onepass=onepass)
            S, Lv, Rv, info = kk.svdsolve(op, u0, 6, "LR", alg)
            rel = np.abs(np.array(S[:6], dtype=np.float64) - truth) / truth
            print(name, "onepass" if onepass else "standard", "converged", info.converged, "numiter", info.numiter, "numops", info.numops,
                  "passes", info.passes, "max rel err", rel.max(), "normres", np.max(info.normres[:6]), f"[{time.time()-t1:.0f} s]", flush=True)
            del Lv, Rv, info
