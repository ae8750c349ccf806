"""Arnoldi factorization — mirror of src/factorizations/arnoldi.jl."""
from __future__ import annotations

import numpy as np

from ..algorithms import Orthogonalizer
from ..dense import hidx
from ..operators import apply
from ..orthonormal import OrthonormalBasis, orthogonalize_
from ..vectors import B200Vec
from . import lanczos as _lz


class ArnoldiIterator:
    """ArnoldiIterator(f, x₀, orth) — arnoldi.jl:103-112."""

    def __init__(self, operator, x0: B200Vec, orth: Orthogonalizer):
        self.operator, self.x0, self.orth = operator, x0, orth


class ArnoldiFactorization:
    """{k, V, H (packed Hessenberg incl. the trailing β), r} — arnoldi.jl:31-36."""

    def __init__(self, k, V: OrthonormalBasis, H: list, r: B200Vec):
        self.k, self.V, self.H, self.r = k, V, H, r

    def __len__(self):
        return self.k

    def basis(self):
        return self.V

    def residual(self):
        return self.r

    def normres(self):
        return abs(self.H[-1])

    def h(self, i: int, j: int) -> float:
        """rayleighquotient(F)[i, j], 1-based — PackedHessenberg(F.H, F.k)."""
        return 0.0 if i > j + 1 else self.H[hidx(i, j)]

    def rayleighquotient(self) -> np.ndarray:
        k = self.k
        Hm = np.zeros((k, k))
        for j in range(1, k + 1):
            for i in range(1, min(j + 1, k) + 1):
                Hm[i - 1, j - 1] = self.H[hidx(i, j)]
        return Hm


def initialize(it: ArnoldiIterator) -> ArnoldiFactorization:
    """initialize(iter) — arnoldi.jl:135-175 (same arithmetic as the Lanczos start)."""
    f = _lz.initialize(_lz.LanczosIterator(it.operator, it.x0, it.orth))
    return ArnoldiFactorization(1, f.V, [f.alphas[0], f.betas[0]], f.r)


def initialize_(it: ArnoldiIterator, state: ArnoldiFactorization) -> ArnoldiFactorization:
    """initialize!(iter, state) — arnoldi.jl:176-198."""
    V = state.V
    while len(V) > 1:
        V.pop()
    V[0] = V[0].scale_(1 / it.x0.norm(), it.x0)
    w = apply(it.operator, V[0])
    r, alpha = orthogonalize_(w, V[0], it.orth)
    beta = r.norm()
    state.k = 1
    state.H = [alpha, beta]
    state.r = r
    return state


def arnoldirecurrence_(operator, V: OrthonormalBasis, h: np.ndarray, orth: Orthogonalizer):
    """arnoldirecurrence!! — arnoldi.jl:239-245.  The norm is a by-product of the fused
    orthogonalisation kernel (same value as a separate norm(r) sweep)."""
    w = apply(operator, V[-1])
    r, h = orthogonalize_(w, V, h, orth)
    return r, orthogonalize_.last_norm


def expand_(it: ArnoldiIterator, state: ArnoldiFactorization) -> ArnoldiFactorization:
    """expand!(iter, state) — arnoldi.jl:199-219."""
    state.k += 1
    k = state.k
    beta = state.normres()
    state.V.push(state.r.scale(1 / beta))      # scale (copy), not scale!! — arnoldi.jl:209
    h = np.empty(k)
    r, beta = arnoldirecurrence_(it.operator, state.V, h, it.orth)
    state.H.extend(h.tolist())
    state.H.append(beta)
    state.r = r
    return state


def shrink_(state: ArnoldiFactorization, k: int) -> ArnoldiFactorization:
    """shrink!(state, k) — arnoldi.jl:220-236."""
    if state.k <= k:
        return state
    V = state.V
    while len(V) > k + 1:
        V.pop()
    r = V.pop()
    del state.H[(k * k + 3 * k) >> 1:]
    state.k = k
    state.r = r.scale_(state.normres())
    return state
