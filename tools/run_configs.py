"""Run the BASELINE.json configs at full size on the GPU(s) and record throughput + size-independent
parity properties.  Single GPU:  python tools/run_configs.py [c1 c2 c3 c4 c5]
Multi GPU (c5 / c2 sharded):  torchrun --nproc-per-node N tools/run_configs.py c5
Writes gpurun_out/configs_<tag>.json."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import krylovkit_jl_b200 as kk  # noqa: E402
from krylovkit_jl_b200 import sharding  # noqa: E402

SEED = 20260923
rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
dist = None
uid = None
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def make_ctx(n_lines, line, ncols, dtype=np.float64):
    """rows = n_lines * line, sharded by whole lines."""
    global uid
    n = n_lines * line
    if world == 1:
        return kk.B200Context(n, ncols, dtype=dtype, device=local), sharding.RowShard(0, 1, 0, n, n)
    import torch
    u = sharding.broadcast_nccl_uid(dist, kk._lib.load())
    sh = sharding.shard_grid_lines(line, n_lines, rank, world)
    return kk.B200Context(sh.n_local, ncols, dtype=dtype, device=local, rank=rank, nranks=world, nccl_uid=u,
                          n_global=n, row_offset=sh.row_offset), sh


def timed(fn, ctx, reps=1):
    fn()
    ctx.lib.b2k_device_sync()
    if dist is not None:
        dist.barrier()
    ctx.lib.b2k_timer_start(ctx.h)
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    ms = C.c_double()
    ctx.lib.b2k_timer_stop(ctx.h, C.byref(ms))
    return out, ms.value / 1000.0 / reps, (time.perf_counter() - t0) / reps


def lanczos_invariants(ctx, op, x0, alg, steps):
    """size-independent parity properties (test/factorize.jl:140-148) evaluated ON DEVICE:
    max |<v_i, v_j> - delta_ij| over a sample, ||r|| = beta, ||A v_j - (T column j)|| small."""
    from krylovkit_jl_b200.factorizations import lanczos as lz
    it = lz.LanczosIterator(op, x0, alg.orth)
    f = lz.initialize(it)
    for _ in range(steps):
        f = lz.expand_(it, f)
    k = f.k
    G = np.zeros((k, k), order="F")
    hs = kk.vectors.handles(f.V.basis)
    ctx.check(ctx.lib.b2k_block_inner(ctx.h, hs, k, hs, k, G.ctypes.data_as(C.POINTER(C.c_double))))
    orth_err = float(np.abs(G - np.eye(k)).max())
    beta_err = abs(f.r.norm() - f.normres()) / f.normres()
    res = []
    for j in (0, k // 2, k - 1):
        w = kk.apply(op, f.V[j])
        w.add_(f.V[j], -f.alphas[j])
        if j > 0:
            w.add_(f.V[j - 1], -f.betas[j - 1])
        if j < k - 1:
            w.add_(f.V[j + 1], -f.betas[j])
        else:
            w.add_(f.r, -1.0)
        res.append(w.norm())
    return {"steps": steps, "max|V'V-I|": orth_err, "|norm(r)-beta|/beta": beta_err,
            "max||A v_j - V T_j - r e_j||": float(max(res))}


def c1():
    """eigsolve(Lanczos, :SR, 4) on the 1e4 x 1e4 5-point Laplacian, krylovdim 30 (reference default MGS2, and CGS2)."""
    from oracle import krylov_oracle as ko
    nx = ny = 100
    out = {}
    ctx = kk.B200Context(nx * ny, 64)
    op = kk.B200CSR.stencil(ctx, nx, ny)
    x0 = ctx.splitmix(SEED)
    A = ko.stencil_matrix(nx, ny)
    xh = ko.splitmix_vector(SEED, nx * ny)
    lam = ko.laplace_eigenvalues(nx, ny)
    for name, orth, oorth in (("mgs2", kk.mgs2, ko.Orth(ko.MGS2)), ("cgs2", kk.cgs2, ko.Orth(ko.CGS2))):
        alg = kk.Lanczos(orth=orth, krylovdim=30, maxiter=100, tol=1e-12, verbosity=0)
        (vals, vecs, info), t_dev, t_wall = timed(lambda: kk.eigsolve(op, x0, 4, "SR", alg), ctx)
        t0 = time.perf_counter()
        ovals, _, oinfo = ko.eigsolve_lanczos(A, xh, 4, "SR", krylovdim=30, maxiter=100, tol=1e-12, orth=oorth)
        t_cpu = time.perf_counter() - t0
        distinct = np.unique(np.round(lam, 12))[:4]
        out[name] = {"numops": info.numops, "numiter": info.numiter, "converged": info.converged,
                     "gpu_s": t_wall, "gpu_it_per_s": info.numops / t_wall,
                     "cpu_oracle_s": t_cpu, "cpu_it_per_s": oinfo["numops"] / t_cpu,
                     "max_rel_diff_vs_oracle": float(np.max(np.abs(vals[:4] - ovals[:4]) / np.abs(ovals[:4]))),
                     "max_rel_diff_vs_closed_form": float(np.max(np.abs(vals[:4] - distinct) / distinct)),
                     "normres": [float(x) for x in info.normres[:4]]}
        del vecs
    ctx.close()
    return out


def c2():
    nx, ny, kd = 4000, 2500, 60
    ctx, sh = make_ctx(ny, nx, kd + 16)
    op = kk.B200CSR.stencil(ctx, nx, ny)
    x0 = ctx.splitmix(SEED)
    out = {}
    for name, orth in (("cgs2", kk.cgs2), ("mgs2", kk.mgs2)):
        alg = kk.Lanczos(orth=orth, krylovdim=kd, maxiter=5, tol=0.0, verbosity=0)
        (vals, vecs, info), t_dev, t_wall = timed(lambda: kk.eigsolve(op, x0, 4, "SR", alg), ctx)
        out[name] = {"numops": info.numops, "s": t_dev, "it_per_s": info.numops / t_dev,
                     "ritz": [float(v) for v in vals[:4]], "normres": [float(v) for v in info.normres[:4]]}
        del vecs
    alg = kk.Lanczos(orth=kk.cgs2, krylovdim=kd, maxiter=1, tol=0.0, verbosity=0)
    out["invariants_cgs2"] = lanczos_invariants(ctx, op, x0, alg, 59)
    ctx.close()
    return out


def c3():
    """linsolve(GMRES, krylovdim=40) on the 1e7 nonsymmetric convection-diffusion CSR, b = A*1, 5 restart cycles."""
    nx, ny, kd = 4000, 2500, 40
    ctx, sh = make_ctx(ny, nx, kd + 16)
    op = kk.B200CSR.stencil(ctx, nx, ny, 1, (4.0, -1.4, -0.6, -1.2, -0.8, 0.0, 0.0))
    ones = ctx.full(1.0)
    b = kk.apply(op, ones)
    out = {}
    for name, orth in (("mgs2", kk.mgs2), ("cgs2", kk.cgs2)):
        alg = kk.GMRES(orth=orth, krylovdim=kd, maxiter=5, tol=1e-300, verbosity=0)
        (x, info), t_dev, t_wall = timed(lambda: kk.linsolve(op, b, None, alg), ctx)
        # b = A x + residual  (test/linsolve.jl:230), evaluated on the device
        chk = kk.apply(op, x)
        chk.add_(info.residual, 1.0).add_(b, -1.0)
        out[name] = {"numops": info.numops, "numiter": info.numiter, "s": t_dev, "it_per_s": info.numops / t_dev,
                     "normres": float(info.normres), "||A x + r - b|| / ||b||": chk.norm() / b.norm(),
                     "||x - 1||_inf proxy ||x-1||/sqrt(n)": float(x.add(ones, -1.0).norm() / np.sqrt(sh.n_global))}
    ctx.close()
    return out


def c4():
    """svdsolve(GKL) on the dense 2e6 x 512 Float32 matrix, 6 triplets, krylovdim 30, tol 1e-5, CGS2."""
    m, n, kd = 2_000_000, 512, 30
    ctx = kk.B200Context(m, kd + 24, dtype=np.float32, device=local)
    sv = ctx.add_space(n, kd + 24, sharded=False)
    op = kk.B200Dense.splitmix(ctx, m, n, SEED, sv)
    u0 = ctx.splitmix(SEED + 1)
    out = {}
    for name, orth in (("mgs2", kk.mgs2), ("cgsr_eta0.75", kk.ClassicalGramSchmidtIR(eta=0.75)), ("cgs2", kk.cgs2)):
        alg = kk.GKL(orth=orth, krylovdim=kd, maxiter=100, tol=1e-5, verbosity=0)
        try:
            (S, Lv, Rv, info), t_dev, t_wall = timed(lambda: kk.svdsolve(op, u0, 6, "LR", alg), ctx)
        except Exception as e:      # single-pass CGS in Float32 can break down on this clustered spectrum
            out[name] = {"failed": f"{type(e).__name__}: {str(e)[:120]}"}
            continue
        # A v = sigma u + r  and  A' u = sigma v  on the device
        res = []
        for i in range(3):
            w = kk.apply_normal(op, Rv[i])
            w.add_(Lv[i], -float(S[i]))
            z = kk.apply_adjoint(op, Lv[i])
            z.add_(Rv[i], -float(S[i]))
            res.append((w.norm() / S[i], z.norm() / S[i]))
        out[name] = {"numops": info.numops, "numiter": info.numiter, "converged": info.converged, "s": t_dev,
                     "it_per_s": info.numops / t_dev, "sigma": [float(s) for s in S[:6]],
                     "rel_residuals(Av-su, A'u-sv)": [[float(a), float(b)] for a, b in res]}
        v0, u1 = Rv[0], Lv[0]
        del Lv, Rv, info
    gemv_bytes = 4.0 * m * n
    xv = ctx.splitmix(7, sv)
    tn = timed(lambda: kk.apply_normal(op, xv), ctx, reps=10)[1]
    tt = timed(lambda: kk.apply_adjoint(op, u0), ctx, reps=10)[1]
    out["gemv_n_GBs"] = gemv_bytes / tn / 1e9
    out["gemv_t_GBs"] = gemv_bytes / tt / 1e9
    ctx.close()
    return out


def c5():
    """eigsolve(Lanczos) on the 8e7 x 8e7 7-point Laplacian (625x500x256), krylovdim 30, CGS2, 3 restart cycles,
    rows sharded by z-planes over the ranks."""
    nx, ny, nz, kd = 625, 500, 256, 30
    ctx, sh = make_ctx(nz, nx * ny, kd + 14)
    op = kk.B200CSR.stencil(ctx, nx, ny, nz, (6.0, -1, -1, -1, -1, -1, -1))
    x0 = ctx.splitmix(SEED)
    alg = kk.Lanczos(orth=kk.cgs2, krylovdim=kd, maxiter=3, tol=0.0, verbosity=0)
    (vals, vecs, info), t_dev, t_wall = timed(lambda: kk.eigsolve(op, x0, 4, "SR", alg), ctx)
    del vecs
    if dist is not None:
        import torch
        tt = torch.tensor([t_dev], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev = float(tt.item())
    out = {"n": nx * ny * nz, "nnz_local": op.nnz, "gpus": world, "numops": info.numops, "s": t_dev,
           "it_per_s": info.numops / t_dev, "ritz": [float(v) for v in vals[:4]]}
    alg1 = kk.Lanczos(orth=kk.cgs2, krylovdim=kd, maxiter=1, tol=0.0, verbosity=0)
    out["invariants"] = lanczos_invariants(ctx, op, x0, alg1, 20)
    ctx.close()
    return out


def cg(nx=4000, ny=2500):
    """linsolve(CG) on the 1e7 5-point Laplacian (SPD), b = A*1, 200 iterations."""
    import importlib
    ls = importlib.import_module("krylovkit_jl_b200.linsolve")
    ctx, sh = make_ctx(ny, nx, 16)
    op = kk.B200CSR.stencil(ctx, nx, ny)
    ones = ctx.full(1.0)
    b = kk.apply(op, ones)
    out = {}
    W = 8.0 * sh.n_local
    spmv_bytes = op.nnz * 12 + 4 * (sh.n_local + 1) + 2 * W
    for name, fused, chain in (("chained_32", True, True), ("fused_step", True, False), ("literal_mirror", False, False)):
        ls.USE_FUSED_CG = fused
        ls.USE_CG_CHAIN = chain
        alg = kk.CG(maxiter=200, tol=1e-300, verbosity=0)
        (x, info), t_dev, t_wall = timed(lambda: kk.linsolve(op, b, None, alg), ctx)
        chk = kk.apply(op, x)
        chk.add_(info.residual, 1.0).add_(b, -1.0)
        it_s = info.numiter / t_dev
        out[name] = {"numiter": info.numiter, "s": t_dev, "it_per_s": it_s, "normres": float(info.normres),
                     "||A x + r - b||/||b||": chk.norm() / b.norm(),
                     "algorithmic_GBs (SpMV + 9W fused / 12W literal)":
                         (spmv_bytes + (9 if fused else 12) * W) * it_s / 1e9}
    ls.USE_FUSED_CG = True
    ls.USE_CG_CHAIN = True
    ctx.close()
    return out


def widened(nx=4000, ny=2500, lnx=2000, lny=2000, lsmr=True):
    """SURVEY §8f rows at the 1e7 scale (1 GPU): BiCGStab and GMRES on the convection-diffusion operator,
    LSMR on a tall sparse least-squares problem, BlockLanczos (p = 4) and Arnoldi eigsolve on the
    Laplacian / convection-diffusion operator, exponentiate (imaginary-time step) on the Laplacian."""
    n = nx * ny
    out = {}
    ctx = kk.B200Context(n, 64)
    W = 8.0 * n
    lap = kk.B200CSR.stencil(ctx, nx, ny)
    cd = kk.B200CSR.stencil(ctx, nx, ny, 1, (4.0, -1.4, -0.6, -1.2, -0.8, 0.0, 0.0))
    ones = ctx.full(1.0)
    spmv_bytes = lap.nnz * 12 + 4 * (n + 1) + 2 * W
    # BiCGStab, 40 iterations (2 operator applications each) on the strictly diagonally dominant variant
    # (centre 5): unpreconditioned BiCGStab stagnates on the centre-4 operator, in the oracle as well
    cd5 = kk.B200CSR.stencil(ctx, nx, ny, 1, (5.0, -1.4, -0.6, -1.2, -0.8, 0.0, 0.0))
    b = ctx.splitmix(SEED + 1)
    import importlib
    ls = importlib.import_module("krylovkit_jl_b200.linsolve")
    alg = kk.BiCGStab(maxiter=40, tol=1e-300, verbosity=0)
    ls.USE_FUSED_BICGSTAB = False
    (x, info), t_lit, _ = timed(lambda: kk.linsolve(cd5, b, None, alg), ctx)
    lit_normres = float(info.normres)
    del x, info
    ls.USE_FUSED_BICGSTAB = True
    (x, info), t_dev, _ = timed(lambda: kk.linsolve(cd5, b, None, alg), ctx)
    chk = kk.apply(cd5, x).add_(b, -1.0)
    out["bicgstab"] = {"numiter": info.numiter, "numops": info.numops, "s": t_dev, "ops_per_s": info.numops / t_dev,
                       "literal_mirror_ops_per_s": info.numops / t_lit, "literal_normres": lit_normres,
                       "normres": float(info.normres), "||A x - b||/||b||": chk.norm() / b.norm(),
                       "algorithmic_GBs (2 SpMV + 17W per iteration)": (2 * spmv_bytes + 17 * W) * info.numiter / t_dev / 1e9}
    del x, info, chk, cd5
    # Arnoldi eigsolve, 3 restart cycles at krylovdim 30
    x0 = ctx.splitmix(SEED)
    alg = kk.Arnoldi(orth=kk.cgs2, krylovdim=30, maxiter=3, tol=0.0, verbosity=0)
    (vals, vecs, info), t_dev, _ = timed(lambda: kk.eigsolve(cd, x0, 2, "LR", alg), ctx)
    out["arnoldi_eigsolve"] = {"numops": info.numops, "numiter": info.numiter, "s": t_dev, "it_per_s": info.numops / t_dev,
                               "ritz": [[float(v.real), float(v.imag)] for v in vals[:2]],
                               "normres": [float(v) for v in info.normres[:2]]}
    del vecs, info
    # BlockLanczos, block of 4, krylovdim 32, 3 restart cycles
    X0 = kk.Block([ctx.splitmix(SEED + i) for i in range(4)])
    for key, fast in (("blocklanczos_p4", False), ("blocklanczos_p4_fast_block", True)):
        alg = kk.BlockLanczos(krylovdim=32, maxiter=3, tol=0.0, verbosity=0, fast_block=fast)
        (vals, vecs, info), t_dev, _ = timed(lambda: kk.eigsolve(lap, X0, 4, "SR", alg), ctx)
        out[key] = {"numops": info.numops, "numiter": info.numiter, "s": t_dev, "it_per_s": info.numops / t_dev,
                    "ritz": [float(v) for v in vals[:4]], "normres": [float(v) for v in info.normres[:4]]}
        del vecs, info
    del X0
    # exponentiate: w = exp(-0.5 A) x0 with Lanczos, krylovdim 30
    alg = kk.Lanczos(orth=kk.cgs2, krylovdim=30, maxiter=10, tol=1e-10, verbosity=0)
    (w, info), t_dev, _ = timed(lambda: kk.exponentiate(lap, -0.5, x0, alg), ctx)
    out["exponentiate"] = {"numops": info.numops, "numiter": info.numiter, "converged": info.converged, "s": t_dev,
                           "it_per_s": info.numops / t_dev, "err_estimate": float(info.normres),
                           "||w||/||x0||": w.norm() / x0.norm()}
    del w, info
    ctx.close()
    if not lsmr:                 # bench.py's GPU arm: the LSMR test matrix below is assembled with the oracle's helper
        return out
    # LSMR: min ||b - A x|| for an 8e6 x 4e6 sparse A = [L; I] (Laplacian stacked on the identity), 50 iterations
    import scipy.sparse as sp
    n = lnx * lny
    m = 2 * n
    ctx = kk.B200Context(m, 12)
    sv = ctx.add_space(n, 24, sharded=False)
    from oracle import krylov_oracle as ko          # operator construction only (host-side test matrix)
    Lh = ko.stencil_matrix(lnx, lny)
    Ah = sp.vstack([Lh, sp.identity(n, format="csr")]).tocsr()
    A = kk.B200CSR.from_scipy(ctx, Ah).with_spaces(sv, 0)
    At = kk.B200CSR.from_scipy(ctx, Ah.T.tocsr()).with_spaces(0, sv)
    bb = ctx.splitmix(SEED)
    alg = kk.LSMR(maxiter=50, krylovdim=8, tol=0.0, verbosity=0)
    (x, info), t_dev, _ = timed(lambda: kk.lssolve((A, At), bb, alg), ctx)
    out["lsmr"] = {"numiter": info.numiter, "numops": info.numops, "s": t_dev, "ops_per_s": info.numops / t_dev,
                   "normres": float(info.normres)}
    ctx.close()
    return out


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5"]
    res = {}
    for w in which:
        res[w] = globals()[w]()
        if rank == 0:
            print(w, json.dumps(res[w]), flush=True)
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        tag = "_".join(which) + f"_n{world}"
        json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"configs_{tag}.json"), "w"), indent=1)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
