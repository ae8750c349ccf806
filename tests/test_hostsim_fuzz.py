"""Randomised agreement of the product's host drivers (run on tests/hostsim.py, literal and fused modes
alternating) with the oracle: same numops / numiter / converged counts and the same results on small random
problems that hit the rare branches — invariant subspaces, eager exits, restarts with tiny Krylov dimensions,
2×2 Schur blocks at the truncation point, rank-deficient start blocks, half-step exits, maxiter quirks.
(A 1000-seed run of this generator is what found the CG `maxiter = 1` discrepancy fixed in round 1.)"""
import warnings

import numpy as np
import pytest
import scipy.sparse as sp

import hostsim
import krylovkit_jl_b200 as kk
from oracle import krylov_oracle as ko

ORTHS = [("cgs", ko.CGS), ("mgs", ko.MGS), ("cgs2", ko.CGS2), ("mgs2", ko.MGS2), ("cgsr", ko.CGSIR), ("mgsr", ko.MGSIR)]


def _orth_pair(rng):
    name, tag = ORTHS[int(rng.integers(0, 6))]
    o = getattr(kk, name)
    return name, o, (ko.Orth(tag, o.eta) if o.is_ir else ko.Orth(tag))


def _sym(B):
    return (B + B.T) / 2


def _counts(info, oinfo, fields=("numops", "numiter", "converged")):
    for f in fields:
        assert getattr(info, f) == oinfo[f], (f, getattr(info, f), oinfo[f])


def _lanczos(rng, seed, ctx, n):
    name, o, ov = _orth_pair(rng)
    d = np.repeat(rng.standard_normal(int(rng.integers(2, 6))), 20)[:n]      # few distinct values: invariant subspaces
    if len(d) < n:
        d = np.concatenate([d, rng.standard_normal(n - len(d))])
    A = sp.diags(d).tocsr() if seed % 4 < 2 else sp.csr_matrix(_sym(rng.standard_normal((n, n))))
    x0 = rng.random(n)
    kd = int(rng.integers(2, min(n, 20) + 1))
    hm = int(rng.integers(1, kd + 1))
    which = ["SR", "LR", "LM"][int(rng.integers(0, 3))]
    alg = kk.Lanczos(orth=o, krylovdim=kd, maxiter=int(rng.integers(1, 8)), tol=1e-9, eager=bool(rng.integers(0, 2)), verbosity=0)
    D, _, info = kk.eigsolve(kk.B200CSR.from_scipy(ctx, A), ctx.from_host(x0), hm, which, alg)
    oD, _, oinfo = ko.eigsolve_lanczos(A, x0, hm, which, krylovdim=kd, maxiter=alg.maxiter, tol=1e-9, orth=ov, eager=alg.eager)
    _counts(info, oinfo)
    if name not in ("cgs", "mgs"):            # without reorthogonalisation ghost values are rounding-chaotic
        assert len(D) == len(oD)
        np.testing.assert_allclose(D, oD, rtol=1e-7, atol=1e-9)


def _arnoldi(rng, seed, ctx, n):
    _, o, ov = _orth_pair(rng)
    A = sp.csr_matrix(rng.random((n, n)) - 0.5)
    x0 = rng.random(n)
    kd = int(rng.integers(3, min(n, 16) + 1))
    hm = int(rng.integers(1, min(kd, 4) + 1))
    which = ["SR", "LR", "LM"][int(rng.integers(0, 3))]
    alg = kk.Arnoldi(orth=o, krylovdim=kd, maxiter=int(rng.integers(1, 6)), tol=1e-9, eager=bool(rng.integers(0, 2)), verbosity=0)
    D, _, info = kk.eigsolve(kk.B200CSR.from_scipy(ctx, A), ctx.from_host(x0), hm, which, alg)
    oD, _, oinfo = ko.eigsolve_arnoldi(A, x0, hm, which, krylovdim=kd, maxiter=alg.maxiter, tol=1e-9, orth=ov, eager=alg.eager)
    _counts(info, oinfo)
    assert len(D) == len(oD)
    np.testing.assert_allclose(D, oD, rtol=1e-6, atol=1e-8)


def _gmres(rng, seed, ctx, n):
    _, o, ov = _orth_pair(rng)
    A = sp.csr_matrix(np.eye(n) * 3 + 0.5 * (rng.random((n, n)) - 0.5))
    b = rng.random(n)
    kd = int(rng.integers(2, 12))
    alg = kk.GMRES(orth=o, krylovdim=kd, maxiter=int(rng.integers(1, 6)), tol=1e-10, verbosity=0)
    a0, a1 = (0.0, 1.0) if seed % 3 else (0.7, 1.3)
    x, info = kk.linsolve(kk.B200CSR.from_scipy(ctx, A), ctx.from_host(b), None, alg, a0, a1)
    ox, oinfo = ko.linsolve_gmres(A, b, krylovdim=kd, maxiter=alg.maxiter, tol=1e-10, orth=ov, a0=a0, a1=a1)
    _counts(info, oinfo, ("numops", "converged"))
    np.testing.assert_allclose(x.to_host(), ox, rtol=1e-7, atol=1e-9)


def _short_recurrences(rng, seed, ctx, n):
    A = sp.csr_matrix(np.eye(n) * 4 + 0.8 * (rng.random((n, n)) - 0.5))
    b = rng.random(n)
    mi = int(rng.integers(1, 30))
    tol = 10.0 ** (-int(rng.integers(3, 12)))
    x, info = kk.linsolve(kk.B200CSR.from_scipy(ctx, A), ctx.from_host(b), None, kk.BiCGStab(maxiter=mi, tol=tol, verbosity=0))
    ox, oinfo = ko.linsolve_bicgstab(A, b, maxiter=mi, tol=tol)
    _counts(info, oinfo)
    np.testing.assert_allclose(x.to_host(), ox, rtol=1e-6, atol=1e-9)
    S = sp.csr_matrix(A @ A.T)
    x, info = kk.linsolve(kk.B200CSR.from_scipy(ctx, S), ctx.from_host(b), None, kk.CG(maxiter=mi, tol=tol, verbosity=0))
    ox, oinfo = ko.linsolve_cg(S, b, maxiter=mi, tol=tol)
    _counts(info, oinfo)
    np.testing.assert_allclose(x.to_host(), ox, rtol=1e-6, atol=1e-9)


def _blocklanczos(rng, seed, ctx, n):
    A = sp.csr_matrix(_sym(rng.standard_normal((n, n))))
    p = int(rng.integers(1, 4))
    X0 = [rng.random(n) for _ in range(p)]
    if p > 1 and seed % 8 == 4:
        X0[1] = 2 * X0[0]                     # linearly dependent start block: block_qr! drops a column
    kd = int(rng.integers(p + 1, min(n, 18) + 1))
    hm = int(rng.integers(1, kd + 1))
    alg = kk.BlockLanczos(krylovdim=kd, maxiter=int(rng.integers(1, 5)), tol=1e-9, qr_tol=1e-10,
                          eager=bool(rng.integers(0, 2)), verbosity=0)
    D, _, info = kk.eigsolve(kk.B200CSR.from_scipy(ctx, A), kk.Block([ctx.from_host(x) for x in X0]), hm, "SR", alg)
    oD, _, oinfo = ko.eigsolve_blocklanczos(A, [x.copy() for x in X0], hm, "SR", krylovdim=kd, maxiter=alg.maxiter,
                                            tol=1e-9, qr_tol=1e-10, eager=alg.eager)
    _counts(info, oinfo)
    assert len(D) == len(oD)
    np.testing.assert_allclose(D, oD, rtol=1e-6, atol=1e-8)


def _lsmr_and_expintegrator(rng, seed, ctx, n):
    _, o, ov = _orth_pair(rng)
    m = n + int(rng.integers(0, 20))
    A = rng.standard_normal((m, n))
    b = rng.random(m)
    K, mi = int(rng.integers(1, 8)), int(rng.integers(1, 9))   # few iterations: see _lsmr_sparse_pair
    lam = float(rng.random() * (seed % 2))
    x, info = kk.lssolve(A, b, kk.LSMR(orth=o, maxiter=mi, krylovdim=K, tol=1e-10, verbosity=0), lam)
    ox, oinfo = ko.lssolve_lsmr(A, b, maxiter=mi, krylovdim=K, tol=1e-10, orth=ov, lam=lam)
    _counts(info, oinfo, ("numiter", "converged"))
    np.testing.assert_allclose(x, ox, rtol=1e-6, atol=1e-9)
    S = sp.csr_matrix(_sym(rng.standard_normal((n, n))) / 2)
    t = float(rng.standard_normal())
    u = tuple(rng.random(n) for _ in range(int(rng.integers(0, 4)) + 1))
    kd, eager = 3 + seed % 9, bool(seed % 2)
    w, info = kk.expintegrator(kk.B200CSR.from_scipy(ctx, S), t, tuple(ctx.from_host(z) for z in u),
                               kk.Lanczos(orth=kk.mgs2, krylovdim=kd, maxiter=50, tol=1e-10, eager=eager, verbosity=0))
    ow, oinfo = ko.expintegrator(S, t, u, "lanczos", ko.Orth(ko.MGS2), krylovdim=kd, maxiter=50, tol=1e-10, eager=eager)
    _counts(info, oinfo)
    np.testing.assert_allclose(w.to_host(), ow, rtol=1e-7, atol=1e-10)


def _svdsolve(rng, seed, ctx, n):
    name, o, ov = _orth_pair(rng)
    m = n + int(rng.integers(0, 12))
    A = rng.standard_normal((m, n))
    u0 = rng.random(m)
    kd = int(rng.integers(2, min(n, 12) + 1))
    hm = int(rng.integers(1, kd + 1))
    which = ["LR", "SR"][int(rng.integers(0, 2))]
    alg = kk.GKL(orth=o, krylovdim=kd, maxiter=int(rng.integers(1, 6)), tol=1e-9, eager=bool(rng.integers(0, 2)), verbosity=0)
    S, _, _, info = kk.svdsolve(A, u0, hm, which, alg)                 # host entry: dense operator, two spaces
    oS, _, _, oinfo = ko.svdsolve_gkl(A, u0, hm, which, krylovdim=kd, maxiter=alg.maxiter, tol=1e-9, orth=ov, eager=alg.eager)
    _counts(info, oinfo)
    if name not in ("cgs", "mgs"):
        assert len(S) == len(oS)
        np.testing.assert_allclose(S, oS, rtol=1e-6, atol=1e-8)


def _schur_and_realeig(rng, seed, ctx, n):
    name, o, ov = _orth_pair(rng)
    A = rng.random((n, n)) - 0.5
    x0 = rng.random(n)
    kd = int(rng.integers(3, min(n, 14) + 1))
    hm = int(rng.integers(1, min(kd, 4) + 1))
    which = ["SR", "LR", "LM"][int(rng.integers(0, 3))]
    alg = kk.Arnoldi(orth=o, krylovdim=kd, maxiter=int(rng.integers(1, 6)), tol=1e-9, eager=bool(rng.integers(0, 2)), verbosity=0)
    kw = dict(krylovdim=kd, maxiter=alg.maxiter, tol=1e-9, orth=ov, eager=alg.eager)
    T, Q, vals, info = kk.schursolve(kk.B200CSR.from_scipy(ctx, sp.csr_matrix(A)), ctx.from_host(x0), hm, which, alg)
    oT, _, ovals, oinfo = ko.schursolve_arnoldi(A, x0, hm, which, **kw)
    _counts(info, oinfo, ("numops", "converged"))
    assert T.shape == oT.shape
    np.testing.assert_allclose(vals, ovals, rtol=1e-6, atol=1e-8)
    Qh = np.column_stack([q.to_host() for q in Q])
    R = np.column_stack([r.to_host() for r in info.residual])
    np.testing.assert_allclose(A @ Qh, Qh @ T + R, atol=1e-8)          # partial Schur relation with residuals
    del Q, info
    As = _sym(A) + np.diag(np.arange(n)) * 0.1                         # real spectrum
    D, _, info = kk.realeigsolve(kk.B200CSR.from_scipy(ctx, sp.csr_matrix(As)), ctx.from_host(x0), hm, which, alg)
    oD, _, oinfo = ko.realeigsolve_arnoldi(As, x0, hm, which, **kw)
    _counts(info, oinfo, ("numops", "converged"))
    assert len(D) == len(oD)
    if name not in ("cgs", "mgs"):
        np.testing.assert_allclose(D, oD, rtol=1e-6, atol=1e-8)


def _gmres_warm_start(rng, seed, ctx, n):
    import importlib
    ls = importlib.import_module("krylovkit_jl_b200.linsolve")
    _, o, ov = _orth_pair(rng)
    A = sp.csr_matrix(np.eye(n) * 3 + 0.5 * (rng.random((n, n)) - 0.5))
    b, xs = rng.random(n), rng.random(n)
    kd = int(rng.integers(2, 10))
    alg = kk.GMRES(orth=o, krylovdim=kd, maxiter=int(rng.integers(1, 5)), tol=1e-10, verbosity=0)
    ls.LITERAL_GIVENS_RESTART = seed % 4 == 2          # gmres.jl:112-117 done literally (k two-column sweeps)
    try:
        x, info = kk.linsolve(kk.B200CSR.from_scipy(ctx, A), ctx.from_host(b), ctx.from_host(xs), alg)
    finally:
        ls.LITERAL_GIVENS_RESTART = False
    ox, oinfo = ko.linsolve_gmres(A, b, xs, krylovdim=kd, maxiter=alg.maxiter, tol=1e-10, orth=ov)
    _counts(info, oinfo, ("numops", "converged"))
    np.testing.assert_allclose(x.to_host(), ox, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(info.residual.to_host(), b - A @ x.to_host(), atol=1e-7)


def _lsmr_sparse_pair(rng, seed, ctx, n):
    """Host entry with a sparse matrix: the (A, Aᵀ) CSR pair in two vector spaces — also when A is square
    (the output space must come from the operator, not from x).  Few iterations: without a long
    reorthogonalisation window LSMR amplifies rounding differences exponentially (SciPy's does, too)."""
    _, o, ov = _orth_pair(rng)
    m = n if seed % 3 == 0 else n + int(rng.integers(1, 15))
    A = (sp.random(m, n, density=0.3, random_state=seed) + sp.eye(m, n)).tocsr()
    b = rng.random(m)
    K, mi = int(rng.integers(1, 6)), int(rng.integers(1, 7))
    x, info = kk.lssolve(A, b, kk.LSMR(orth=o, maxiter=mi, krylovdim=K, tol=1e-10, verbosity=0))
    ox, oinfo = ko.lssolve_lsmr(A.toarray(), b, maxiter=mi, krylovdim=K, tol=1e-10, orth=ov)
    _counts(info, oinfo, ("numiter", "converged"))
    np.testing.assert_allclose(x, ox, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(info.residual, b - A @ x, atol=1e-8)


KINDS = [_lanczos, _arnoldi, _gmres, _short_recurrences, _blocklanczos, _lsmr_and_expintegrator,
         _svdsolve, _schur_and_realeig, _gmres_warm_start, _lsmr_sparse_pair]


@pytest.mark.parametrize("seed", range(160))
def test_driver_agrees_with_oracle(seed):
    warnings.simplefilter("ignore")
    rng = np.random.default_rng(seed)
    with hostsim.installed(fused=bool(seed % 2)):
        n = int(rng.integers(6, 60))
        ctx = kk.B200Context(n, 600)
        KINDS[seed % len(KINDS)](rng, seed, ctx, n)
        ctx.close()


def _blocklanczos_fast_block(rng, seed, ctx, n):
    """The flagged block mode (SpMM + BCGS2 + CholeskyQR2, falling back to the reference block_qr! when a pivot is
    too small): a different orthogonalisation, the same counts and Ritz values as the reference algorithm."""
    A = sp.csr_matrix(_sym(rng.standard_normal((n, n))))
    p = int(rng.integers(1, 5))
    X0 = [rng.random(n) for _ in range(p)]
    if p > 1 and seed % 8 == 4:
        X0[1] = 2 * X0[0]
    kd = int(rng.integers(p + 1, min(n, 18) + 1))
    hm = int(rng.integers(1, kd + 1))
    alg = kk.BlockLanczos(krylovdim=kd, maxiter=int(rng.integers(1, 5)), tol=1e-9, qr_tol=1e-10,
                          eager=bool(rng.integers(0, 2)), verbosity=0, fast_block=True)
    D, V, info = kk.eigsolve(kk.B200CSR.from_scipy(ctx, A), kk.Block([ctx.from_host(x) for x in X0]), hm, "SR", alg)
    oD, _, oinfo = ko.eigsolve_blocklanczos(A, [x.copy() for x in X0], hm, "SR", krylovdim=kd, maxiter=alg.maxiter,
                                            tol=1e-9, qr_tol=1e-10, eager=alg.eager)
    _counts(info, oinfo)
    assert len(D) == len(oD)
    np.testing.assert_allclose(D, oD, rtol=1e-6, atol=1e-8)
    for i in range(min(2, len(V))):                                    # A v = λ v + r
        assert np.linalg.norm(A @ V[i].to_host() - D[i] * V[i].to_host() - info.residual[i].to_host()) < 1e-8
    del V, info


def _blocked_mgs2(rng, seed, ctx, n):
    """The flagged blocked form of the reference-default orthogonalizer (kk.mgs2b) in the Lanczos and GMRES drivers
    against the oracle's literal ModifiedGramSchmidt2."""
    A = sp.csr_matrix(_sym(rng.standard_normal((n, n))))
    x0 = rng.random(n)
    kd = int(rng.integers(2, min(n, 20) + 1))
    hm = int(rng.integers(1, kd + 1))
    which = ["SR", "LR", "LM"][int(rng.integers(0, 3))]
    mi, eager = int(rng.integers(1, 8)), bool(rng.integers(0, 2))
    alg = kk.Lanczos(orth=kk.mgs2b, krylovdim=kd, maxiter=mi, tol=1e-9, eager=eager, verbosity=0)
    D, _, info = kk.eigsolve(kk.B200CSR.from_scipy(ctx, A), ctx.from_host(x0), hm, which, alg)
    oD, _, oinfo = ko.eigsolve_lanczos(A, x0, hm, which, krylovdim=kd, maxiter=mi, tol=1e-9, orth=ko.Orth(ko.MGS2), eager=eager)
    _counts(info, oinfo)
    np.testing.assert_allclose(D, oD, rtol=1e-7, atol=1e-9)
    G = sp.csr_matrix(np.eye(n) * 3 + 0.5 * (rng.random((n, n)) - 0.5))
    b = rng.random(n)
    kd, mi = int(rng.integers(2, 12)), int(rng.integers(1, 6))
    x, info = kk.linsolve(kk.B200CSR.from_scipy(ctx, G), ctx.from_host(b), None,
                          kk.GMRES(orth=kk.mgs2b, krylovdim=kd, maxiter=mi, tol=1e-10, verbosity=0))
    ox, oinfo = ko.linsolve_gmres(G, b, krylovdim=kd, maxiter=mi, tol=1e-10, orth=ko.Orth(ko.MGS2))
    _counts(info, oinfo, ("numops", "converged"))
    np.testing.assert_allclose(x.to_host(), ox, rtol=1e-7, atol=1e-9)


def _gkl_onepass(rng, seed, ctx, n):
    """The flagged one-pass GKL step (GKL(onepass=True): A'u recovered from z = A'(A v), with the error estimate that
    falls back to a direct product) through svdsolve on random problems — all six orthogonalizers, :LR / :SR, eager or
    not, few restarts, start vectors inside and outside range(A), rank-deficient A: the same counts and singular values
    as the oracle's two-pass svdsolve, and never more passes over A than the two-pass step."""
    name, o, ov = _orth_pair(rng)
    m = n + int(rng.integers(0, 40))
    A = rng.standard_normal((m, n))
    if seed % 5 == 0:
        A[:, -1] = A[:, 0]                                       # rank deficient
    u0 = rng.random(m) if seed % 3 else A @ rng.random(n)         # every third start vector lies in range(A)
    kd = int(rng.integers(2, min(n, 12) + 1))
    hm = int(rng.integers(1, kd + 1))
    which = ["LR", "SR"][int(rng.integers(0, 2))]
    alg = kk.GKL(orth=o, krylovdim=kd, maxiter=int(rng.integers(1, 6)), tol=1e-9, eager=bool(rng.integers(0, 2)), verbosity=0,
                 onepass=True)
    S, Lv, Rv, info = kk.svdsolve(A, u0, hm, which, alg)
    oS, _, _, oinfo = ko.svdsolve_gkl(A, u0, hm, which, krylovdim=kd, maxiter=alg.maxiter, tol=1e-9, orth=ov, eager=alg.eager)
    # (without reorthogonalisation — cgs / mgs — a residual at rounding level makes the count of "converged" values
    # noise, in the two-pass step as much as here)
    _counts(info, oinfo, ("numops", "numiter") if name in ("cgs", "mgs") else ("numops", "numiter", "converged"))
    assert info.numops // 2 + 1 <= info.passes <= info.numops
    if name not in ("cgs", "mgs"):
        assert len(S) == len(oS)
        np.testing.assert_allclose(S, oS, rtol=1e-6, atol=1e-8)
        for i in range(min(2, len(S))):                          # A v = s u + r, A'u = s v (u is arbitrary for s = 0)
            if S[i] < 1e-8 * max(S):
                continue
            assert np.linalg.norm(A.T @ Lv[i] - S[i] * Rv[i]) < 1e-6 * max(1.0, S[0])
            assert np.linalg.norm(A @ Rv[i] - S[i] * Lv[i] - info.residual[i]) < 1e-6 * max(1.0, S[0])


@pytest.mark.parametrize("seed", range(24))
@pytest.mark.parametrize("kind", [_blocklanczos_fast_block, _blocked_mgs2, _gkl_onepass], ids=["fast_block", "mgs2b", "gkl_onepass"])
def test_flagged_modes_agree_with_oracle(kind, seed):
    """(300 seeds of each ran clean when these were added.)"""
    warnings.simplefilter("ignore")
    rng = np.random.default_rng(10_000 + seed)
    with hostsim.installed(fused=True):
        n = int(rng.integers(8, 60))
        ctx = kk.B200Context(n, 600)
        kind(rng, 10_000 + seed, ctx, n)
        ctx.close()
