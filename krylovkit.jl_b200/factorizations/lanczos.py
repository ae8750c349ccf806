"""Lanczos factorization — mirror of src/factorizations/lanczos.jl.

`expand_` has two implementations with identical semantics:
  * the literal mirror (apply + VectorInterface calls, one C-ABI call per reference op);
  * the fused C-ABI call b2k_lanczos_expand (one host round trip per step) used when the
    operator is a device CSR matrix.  tests/test_gpu_factorizations.py checks they agree.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from .. import _lib as L
from ..algorithms import Orthogonalizer
from ..operators import B200CSR, apply
from ..orthonormal import OrthonormalBasis, orthogonalize_
from ..vectors import B200Vec, handles

USE_FUSED_EXPAND = True
EPS = float(np.finfo(np.float64).eps)


class LanczosIterator:
    """LanczosIterator(f, x₀, orth, keepvecs) — lanczos.jl:130-146."""

    def __init__(self, operator, x0: B200Vec, orth: Orthogonalizer, keepvecs: bool = True):
        if not keepvecs and (orth.is_reorth2 or orth.is_ir):
            raise ValueError("Cannot use reorthogonalization without keeping all Krylov vectors")
        self.operator, self.x0, self.orth, self.keepvecs = operator, x0, orth, keepvecs


class LanczosFactorization:
    """{k, V, αs, βs, r} — lanczos.jl:31-37."""

    def __init__(self, k, V: OrthonormalBasis, alphas, betas, r: B200Vec):
        self.k, self.V, self.alphas, self.betas, self.r = k, V, alphas, betas, r

    def __len__(self):
        return self.k

    def basis(self):
        if len(self.V) != self.k:
            raise RuntimeError("Not keeping vectors during Lanczos factorization")
        return self.V

    def rayleighquotient(self):
        """SymTridiagonal(αs, βs) as (dv, ev)."""
        return np.array(self.alphas), np.array(self.betas[: self.k - 1])

    def residual(self):
        return self.r

    def normres(self):
        return self.betas[-1]


def _eps(v: B200Vec):
    return float(np.finfo(v.ctx.np_dtype).eps)


def initialize(it: LanczosIterator) -> LanczosFactorization:
    """initialize(iter) — lanczos.jl:180-222."""
    x0 = it.x0
    beta0 = x0.norm()
    if beta0 == 0:
        raise ValueError("initial vector should not have norm zero")
    Ax0 = apply(it.operator, x0)
    alpha = x0.inner(Ax0) / (beta0 * beta0)
    v = x0.scale(1 / beta0)                    # add!!(scale(Ax₀, 0), x₀, 1/β₀)
    r = Ax0.scale_(1 / beta0)
    betaold = r.norm()
    r = r.add_(v, -alpha)
    beta = r.norm()
    if it.orth.is_reorth2:
        dalpha = v.inner(r)
        alpha += dalpha
        r = r.add_(v, -dalpha)
        beta = r.norm()
    elif it.orth.is_ir:
        while _eps(r) < beta < it.orth.eta * betaold:
            betaold = beta
            dalpha = v.inner(r)
            alpha += dalpha
            r = r.add_(v, -dalpha)
            beta = r.norm()
    return LanczosFactorization(1, OrthonormalBasis([v]), [alpha], [beta], r)


def initialize_(it: LanczosIterator, state: LanczosFactorization) -> LanczosFactorization:
    """initialize!(iter, state) — lanczos.jl:223-249."""
    V = state.V
    while len(V) > 1:
        V.pop()
    state.alphas.clear()
    state.betas.clear()
    V[0] = V[0].scale_(1 / it.x0.norm(), it.x0)
    w = apply(it.operator, V[0])
    r, alpha = orthogonalize_(w, V[0], it.orth)
    beta = r.norm()
    state.k = 1
    state.alphas.append(alpha)
    state.betas.append(beta)
    state.r = r
    return state


def lanczosrecurrence(operator, V: OrthonormalBasis, beta: float, orth: Orthogonalizer):
    """lanczosrecurrence ×6 — lanczos.jl:295-376 (literal mirror)."""
    t = orth.tag
    v = V[-1]
    w = apply(operator, v)
    if t == L.CGS:
        alpha = v.inner(w)
        w = w.add_(V[-2], -beta)
        w = w.add_(v, -alpha)
        return w, alpha, w.norm()
    if t == L.MGS:
        w = w.add_(V[-2], -beta)
        alpha = v.inner(w)
        w = w.add_(v, -alpha)
        return w, alpha, w.norm()
    if t == L.CGS2:
        alpha = v.inner(w)
        w = w.add_(V[-2], -beta)
        w = w.add_(v, -alpha)
        from ..algorithms import cgs
        w, s = orthogonalize_(w, V, cgs)
        alpha += s[len(V) - 1]
        return w, alpha, w.norm()
    if t == L.MGS2:
        from ..algorithms import mgs
        w = w.add_(V[-2], -beta)
        w, alpha = orthogonalize_(w, v, mgs)
        s = alpha
        for q in V:
            w, s = orthogonalize_(w, q, mgs)
        alpha += s
        return w, alpha, w.norm()
    if t == L.MGS2B:
        # flagged: the reference's MGS2 recurrence with its second sweep over V applied as one classical block
        from ..algorithms import cgs, mgs
        w = w.add_(V[-2], -beta)
        w, alpha = orthogonalize_(w, v, mgs)
        w, s = orthogonalize_(w, V, cgs)
        alpha += s[len(V) - 1]
        return w, alpha, w.norm()
    if t == L.CGSIR:
        from ..algorithms import cgs
        alpha = v.inner(w)
        w = w.add_(V[-2], -beta)
        w = w.add_(v, -alpha)
        ab2 = alpha * alpha + beta * beta
        beta = w.norm()
        nold = math.sqrt(beta * beta + ab2)
        while _eps(w) < beta < orth.eta * nold:
            nold = beta
            w, s = orthogonalize_(w, V, cgs)
            alpha += s[len(V) - 1]
            beta = w.norm()
        return w, alpha, beta
    if t == L.MGSIR:
        from ..algorithms import mgs
        w = w.add_(V[-2], -beta)
        w, alpha = orthogonalize_(w, v, mgs)
        ab2 = alpha * alpha + beta * beta
        beta = w.norm()
        nold = math.sqrt(beta * beta + ab2)
        while _eps(w) < beta < orth.eta * nold:
            nold = beta
            s = 0.0
            for q in V:
                w, s = orthogonalize_(w, q, mgs)
            alpha += s
            beta = w.norm()
        return w, alpha, beta
    raise ValueError(f"unknown orthogonalizer {orth}")


def expand_(it: LanczosIterator, state: LanczosFactorization, fused: bool | None = None):
    """expand!(iter, state) — lanczos.jl:250-272."""
    betaold = state.normres()
    V, r = state.V, state.r
    use_fused = USE_FUSED_EXPAND if fused is None else fused
    if use_fused and isinstance(it.operator, B200CSR) and it.keepvecs:
        ctx = r.ctx
        w = ctx.empty(r.space)
        cols = handles(list(V.basis) + [r])
        a, b = C.c_double(), C.c_double()
        ctx.check(ctx.lib.b2k_lanczos_expand(ctx.h, it.operator.h, cols, len(V), r.handle, w.handle,
                                             betaold, it.orth.tag, it.orth.eta, C.byref(a), C.byref(b)))
        V.push(r)                              # the residual's storage became the new basis vector
        alpha, beta, r = a.value, b.value, w
    else:
        V.push(r.scale_(1 / betaold))
        r, alpha, beta = lanczosrecurrence(it.operator, V, betaold, it.orth)
    state.alphas.append(alpha)
    state.betas.append(beta)
    if not it.keepvecs:
        V.popfirst()
    state.k += 1
    state.r = r
    return state


def expand_many_(it: LanczosIterator, state: LanczosFactorization, nsteps: int, tol: float) -> int:
    """Up to `nsteps` consecutive expand! steps in one C-ABI call (b2k_lanczos_expand_many); stops
    early once normres <= tol.  Returns the number of steps done.  Falls back to a loop of
    expand_ for operators that are not device CSR matrices."""
    if nsteps <= 0:
        return 0
    if not (USE_FUSED_EXPAND and isinstance(it.operator, B200CSR) and it.keepvecs):
        done = 0
        for _ in range(nsteps):
            expand_(it, state)
            done += 1
            if state.normres() <= tol:
                break
        return done
    V, r = state.V, state.r
    ctx = r.ctx
    k = len(V)
    cols = (L.c_vec * (k + nsteps + 1))()
    for i, v in enumerate(V.basis):
        cols[i] = v.handle
    cols[k] = r.handle
    al = (C.c_double * nsteps)()
    be = (C.c_double * nsteps)()
    done, rout = C.c_int32(), L.c_vec()
    status = ctx.lib.b2k_lanczos_expand_many(ctx.h, it.operator.h, cols, k, nsteps, state.normres(), tol,
                                             it.orth.tag, it.orth.eta, al, be, C.byref(done), C.byref(rout))
    # commit the completed steps BEFORE raising: the library has already turned the residual into a basis
    # vector and allocated columns for them, so the factorization must stay consistent for a caller that
    # catches the error (e.g. a slab that ran out of columns at step i > 0)
    d = done.value
    if d > 0:
        # the library returns the d new basis vectors in cols[k : k + d] and the residual in r_out.  On the
        # literal path the old residual's column IS the first of them (lanczos.jl:257); on the device-chained
        # path the normalised vector was written to a column of its own and r's column has been recycled.
        new = [int(cols[k + i]) for i in range(d)]
        if new[0] == r.handle:
            V.push(r)
        else:
            r.disown()                              # released (and possibly reused) inside the library
            V.push(B200Vec(ctx, new[0]))
        for h in new[1:]:
            V.push(B200Vec(ctx, h))                 # columns allocated by the library
        state.r = B200Vec(ctx, rout.value)
        state.alphas.extend(al[:d])
        state.betas.extend(be[:d])
        state.k += d
    ctx.check(status)
    return d


def shrink_(state: LanczosFactorization, k: int):
    """shrink!(state, k) — lanczos.jl:273-291."""
    if state.k != len(state.V):
        raise RuntimeError("we cannot shrink LanczosFactorization without keeping Lanczos vectors")
    if state.k <= k:
        return state
    V = state.V
    while len(V) > k + 1:
        V.pop()
    r = V.pop()
    del state.alphas[k:]
    del state.betas[k:]
    state.k = k
    state.r = r.scale_(state.normres())
    return state
