"""lssolve with LSMR — mirror of src/lssolve/lsmr.jl (front end: src/lssolve/lssolve.jl).

min_x ‖b − A x‖² + λ²‖x‖² by Golub-Kahan bidiagonalisation; every vector operation is a
VectorInterface call on device vectors, the scalar rotations stay on the host.  The operator is
anything `apply_normal` / `apply_adjoint` accept: a B200Dense, a pair (A, Aᵀ) of B200CSR
operators (rectangular ones carry `space_in` / `space_out`), or a callable f(x, flag).
"""
from __future__ import annotations

import math
import warnings

import numpy as np

from .algorithms import ConvergenceInfo, LSMR, WARN_LEVEL
from .operators import B200CSR, B200Dense, apply_adjoint, apply_normal
from .orthonormal import OrthonormalBasis, orthogonalize_
from .vectors import B200Context, B200Vec


def lssolve(A, b, alg: LSMR | None = None, lam: float = 0.0, atol: float | None = None,
            rtol: float | None = None, **kwargs):
    """lssolve(A, b, alg::LSMR, λ).  Host entry: A = numpy m×n array or scipy sparse matrix and
    b = numpy vector -> uploaded, solved, downloaded.  tol = max(atol, rtol*‖Aᴴb‖) when atol/rtol
    are given (lssolve.jl:118-124)."""
    if alg is None:
        alg = LSMR(**kwargs)
    if isinstance(b, B200Vec):
        if atol is not None or rtol is not None:
            alg = _retol(alg, max(atol or 0.0, (rtol or 0.0) * apply_adjoint(A, b).norm()))
        return _lsmr(A, b, alg, lam)
    import scipy.sparse as sp
    b = np.asarray(b)
    m, n = A.shape
    dtype = np.float32 if b.dtype == np.float32 else np.float64
    ctx = B200Context(m, 12, dtype=dtype)
    try:
        sv = ctx.add_space(n, alg.krylovdim + 10, sharded=False)
        if sp.issparse(A):
            op = (B200CSR.from_scipy(ctx, A).with_spaces(sv, 0),
                  B200CSR.from_scipy(ctx, A.T.tocsr()).with_spaces(0, sv))
        else:
            op = B200Dense.from_host(ctx, np.asarray(A), sv)
        x, info = lssolve(op, ctx.from_host(b), alg, lam, atol, rtol)
        info.residual = info.residual.to_host()
        return x.to_host(), info
    finally:
        ctx.close()


def _retol(alg: LSMR, tol: float) -> LSMR:
    return LSMR(orth=alg.orth, maxiter=alg.maxiter, krylovdim=alg.krylovdim, tol=tol, verbosity=alg.verbosity)


def _lsmr(operator, b: B200Vec, alg: LSMR, lam: float):
    """lssolve(operator, b, alg::LSMR, λ) — lsmr.jl:1-162."""
    u = b.copy()
    v = apply_adjoint(operator, b)
    beta = u.norm()
    u = u.scale_(1 / beta)
    v = v.scale_(1 / beta)
    alpha = v.norm()
    v = v.scale_(1 / alpha)

    V = OrthonormalBasis([v])
    K = alg.krylovdim
    Vv = np.zeros(K)                       # storage for the reorthogonalisation coefficients

    alphabar, zetabar = alpha, alpha * beta
    rho, theta, rhobar, cbar, sbar = 1.0, 0.0, 1.0, 1.0, 0.0
    abszetabar = abs(zetabar)

    x = v.zerovector()
    h = v.copy()                           # a copy: v itself sits in the ring and gets replaced
    hbar = v.zerovector()
    r = u.scale(beta)
    Ah = u.zerovector()
    Ahbar = u.zerovector()

    numiter, numops = 0, 1
    maxiter, tol = alg.maxiter, alg.tol
    if abszetabar < tol:
        return x, ConvergenceInfo(1, r, abszetabar, numiter, numops)

    while True:
        numiter += 1
        Av = apply_normal(operator, v)
        numops += 1
        Ah = Ah.add_(Av, 1.0, -theta / rho)

        # β₊ u₊ = A v − α u
        u = Av.add_(u, -alpha)
        del Av
        beta = u.norm()
        if beta > tol:
            u = u.scale_(1 / beta)
            # α₊ v₊ = Aᴴ u₊ − β₊ v
            v = apply_adjoint(operator, u).add_(v, -beta)
            numops += 1
            if K > 1:                      # reorthogonalise against the ring, in slot order
                v, _ = orthogonalize_(v, V, Vv[:min(K, numiter)], alg.orth)
            alpha = v.norm()
            if alpha > tol:
                v = v.scale_(1 / alpha)
                if numiter < K:
                    V.push(v)
                else:
                    V[numiter % K] = v     # mod1(numiter + 1, K) in 1-based terms

        # rotation P̂ (folds the regularisation λ into ᾱ)
        alphahat = math.hypot(alphabar, lam)
        # rotation P: B → R
        rhoold = rho
        rho = math.hypot(alphahat, beta)
        c, s = alphahat / rho, beta / rho
        theta = s * alpha
        alphabar = c * alpha
        # rotation P̄: Rᵀ → R̄
        rhobarold = rhobar
        thetabar = sbar * rho
        cbarrho = cbar * rho
        rhobar = math.hypot(cbarrho, theta)
        cbar = cbarrho / rhobar
        sbar = theta / rhobar
        zeta = cbar * zetabar
        zetabar = -sbar * zetabar

        g = -thetabar * rho / (rhoold * rhobarold)
        hbar = hbar.add_(h, 1.0, g)        # h̄ ← h + g h̄
        Ahbar = Ahbar.add_(Ah, 1.0, g)
        x = x.add_(hbar, zeta / (rho * rhobar))
        r = r.add_(Ahbar, -zeta / (rho * rhobar))
        h = h.add_(v, 1.0, -theta / rho)   # Ah catches up at the top of the next iteration

        abszetabar = abs(zetabar)
        if abszetabar <= tol:
            return x, ConvergenceInfo(1, r, abszetabar, numiter, numops)
        if numiter >= maxiter:
            if alg.verbosity >= WARN_LEVEL:
                warnings.warn(f"LSMR lssolve stopped without converging after {numiter} iterations: "
                              f"normres = {abszetabar}, numops = {numops}")
            return x, ConvergenceInfo(0, r, abszetabar, numiter, numops)
