"""Golub-Kahan-Lanczos bidiagonalisation — mirror of src/factorizations/gkl.jl."""
from __future__ import annotations

import math

import numpy as np

from .. import _lib as L
from ..algorithms import Orthogonalizer, cgs, mgs
from ..operators import apply_adjoint, apply_normal
from ..orthonormal import OrthonormalBasis, orthogonalize_
from ..vectors import B200Vec


class GKLIterator:
    """GKLIterator(f, u₀, orth, keepvecs) — gkl.jl:123-139.  u₀ lives in the codomain."""

    def __init__(self, operator, u0: B200Vec, orth: Orthogonalizer, keepvecs: bool = True):
        if not keepvecs and (orth.is_reorth2 or orth.is_ir):
            raise ValueError("Cannot use reorthogonalization without keeping all Krylov vectors")
        self.operator, self.u0, self.orth, self.keepvecs = operator, u0, orth, keepvecs


class GKLFactorization:
    """{k, U, V, αs, βs, r} — gkl.jl:31-38; rayleighquotient = Bidiagonal(αs, βs, :L)."""

    def __init__(self, k, U: OrthonormalBasis, V: OrthonormalBasis, alphas, betas, r: B200Vec):
        self.k, self.U, self.V, self.alphas, self.betas, self.r = k, U, V, alphas, betas, r

    def __len__(self):
        return self.k

    def basis(self, which: str):
        return self.U if which == "U" else self.V

    def residual(self):
        return self.r

    def normres(self):
        return self.betas[-1]


def _eps(v):
    return float(np.finfo(v.ctx.np_dtype).eps)


def initialize(it: GKLIterator) -> GKLFactorization:
    """initialize(iter) — gkl.jl:183-215."""
    u0 = it.u0
    beta0 = u0.norm()
    if beta0 == 0:
        raise ValueError("initial vector should not have norm zero")
    v0 = apply_adjoint(it.operator, u0)
    alpha = v0.norm() / beta0
    Av0 = apply_normal(it.operator, v0)
    alpha2 = u0.inner(Av0) / (beta0 * beta0)
    if not math.isclose(alpha2, alpha * alpha, rel_tol=math.sqrt(_eps(u0))):
        raise ValueError("operator and its adjoint are not compatible")
    u = u0.scale(1 / beta0)
    v = v0.scale(1 / (alpha * beta0))
    r = Av0.scale_(1 / (alpha * beta0))
    r = r.add_(u, -alpha)
    beta = r.norm()
    v0.free()
    return GKLFactorization(1, OrthonormalBasis([u]), OrthonormalBasis([v]), [alpha], [beta], r)


def gklrecurrence(operator, U: OrthonormalBasis, V: OrthonormalBasis, beta: float,
                  orth: Orthogonalizer):
    """gklrecurrence ×5 — gkl.jl:294-404."""
    t = orth.tag
    u = U[-1]
    v = apply_adjoint(operator, u)
    v = v.add_(V[-1], -beta)
    if t == L.MGS2:
        for q in V:
            v, _ = orthogonalize_(v, q, mgs)
    elif t == L.MGS2B:                          # flagged: the same sweep as one classical block
        v, _ = orthogonalize_(v, V, cgs)
    alpha = v.norm()
    if t in (L.CGSIR, L.MGSIR):
        nold = math.sqrt(alpha * alpha + beta * beta)
        while (alpha < orth.eta * nold) if t == L.CGSIR else (_eps(v) < alpha < orth.eta * nold):
            nold = alpha
            if t == L.CGSIR:
                v, _ = orthogonalize_(v, V, cgs)
            else:
                for q in V:
                    v, _ = orthogonalize_(v, q, mgs)
            alpha = v.norm()
    v = v.scale_(1 / alpha)

    r = apply_normal(operator, v)
    r = r.add_(u, -alpha)
    beta_known = None
    if t == L.CGS2:
        r, _ = orthogonalize_(r, U, cgs)        # only the long (U) side is reorthogonalised, :320
        beta_known = orthogonalize_.last_norm
    elif t == L.MGS2:
        for q in U:
            r, _ = orthogonalize_(r, q, mgs)
    elif t == L.MGS2B:
        r, _ = orthogonalize_(r, U, cgs)
        beta_known = orthogonalize_.last_norm
    beta = beta_known if beta_known is not None else r.norm()
    if t in (L.CGSIR, L.MGSIR):
        nold = math.sqrt(alpha * alpha + beta * beta)
        while _eps(r) < beta < orth.eta * nold:
            nold = beta
            if t == L.CGSIR:
                r, _ = orthogonalize_(r, U, cgs)
            else:
                for q in U:
                    r, _ = orthogonalize_(r, q, mgs)
            beta = r.norm()
    return v, r, alpha, beta


def expand_(it: GKLIterator, state: GKLFactorization) -> GKLFactorization:
    """expand!(iter, state) — gkl.jl:246-269."""
    betaold = state.normres()
    U, V, r = state.U, state.V, state.r
    U.push(r.scale_(1 / betaold))
    v, r, alpha, beta = gklrecurrence(it.operator, U, V, betaold, it.orth)
    V.push(v)
    state.alphas.append(alpha)
    state.betas.append(beta)
    state.k += 1
    state.r = r
    return state


def shrink_(state: GKLFactorization, k: int) -> GKLFactorization:
    """shrink!(state, k) — gkl.jl:270-291."""
    if state.k != len(state.V):
        raise RuntimeError("we cannot shrink GKLFactorization without keeping vectors")
    if state.k <= k:
        return state
    U, V = state.U, state.V
    while len(V) > k + 1:
        U.pop()
        V.pop()
    V.pop()
    r = U.pop()
    del state.alphas[k:]
    del state.betas[k:]
    state.k = k
    state.r = r.scale_(state.normres())
    return state
