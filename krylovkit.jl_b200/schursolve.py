"""Arnoldi / Krylov-Schur drivers — mirror of src/eigsolve/arnoldi.jl.

`schursolve` returns a partial real Schur decomposition, `eigsolve(…, ::Arnoldi)` the eigenpairs of a
general (non-symmetric) real operator.  All n-length work is the Arnoldi expansion, one
basistransform! per restart and the final Ritz-vector combinations; the projected problem is handled
on the host as in the reference.  Complex eigenvectors of a real problem are returned as
`ComplexVec(re, im)` pairs of device vectors (the library has no complex dtype).
"""
from __future__ import annotations

import warnings

import numpy as np

from .algorithms import Arnoldi, ConvergenceInfo, WARN_LEVEL
from .dense import (eigsort_complex, hidx, hschur, permuteschur, restore_arnoldi_form, schur2eigvals,
                    schur2eigvecs, schur2realeigvecs)
from .factorizations import arnoldi as ar
from .orthonormal import basistransform_
from .vectors import B200Vec


class ComplexVec:
    """re + i·im with both parts resident on the device."""

    __slots__ = ("re", "im")

    def __init__(self, re: B200Vec, im: B200Vec):
        self.re, self.im = re, im

    def to_host(self) -> np.ndarray:
        return self.re.to_host().astype(np.complex128) + 1j * self.im.to_host()

    def norm(self) -> float:
        return float(np.hypot(self.re.norm(), self.im.norm()))


def _schursolve(A, x0: B200Vec, howmany: int, which: str, alg: Arnoldi):
    """_schursolve — eigsolve/arnoldi.jl:351-452."""
    krylovdim, maxiter = alg.krylovdim, alg.maxiter
    if howmany > krylovdim:
        raise ValueError(f"krylov dimension {krylovdim} too small to compute {howmany} eigenvalues")
    sort = eigsort_complex(which)
    numiter = 1
    it = ar.ArnoldiIterator(A, x0, alg.orth)
    fact = ar.initialize(it)
    numops = 1
    tol = alg.tol
    converged = 0
    T = U = f = None
    while True:
        beta = fact.normres()
        K = len(fact)
        if beta <= tol and K < howmany and alg.verbosity >= WARN_LEVEL:
            warnings.warn(f"Invariant subspace of dimension {K} (up to requested tolerance `tol = {tol}`), "
                          f"which is smaller than the number of requested eigenvalues (i.e. `howmany == {howmany}`).")
        if K == krylovdim or beta <= tol or (alg.eager and K >= howmany):
            T, U, values = hschur(fact.rayleighquotient())
            T, U, values = permuteschur(T, U, sort(values))
            f = U[K - 1, :] * beta
            converged = 0
            while converged < K and abs(f[converged]) <= tol:
                converged += 1
            if 0 < converged < K and T[converged, converged - 1] != 0:
                converged -= 1                       # do not count half of a conjugate pair
            if converged >= howmany or beta <= tol:
                break
        if K < krylovdim:
            fact = ar.expand_(it, fact)
            numops += 1
        else:
            if numiter == maxiter:
                break
            keep = (3 * krylovdim + 2 * converged) // 5
            H = np.array(T)
            if H[keep, keep - 1] != 0:
                # in the middle of a 2×2 block: shrink by one more, but never to zero
                if keep > 1:
                    keep -= 1
                else:
                    keep += 1
                    if krylovdim == 2:
                        if alg.verbosity >= WARN_LEVEL:
                            warnings.warn("Arnoldi iteration got stuck in a 2x2 subspace that cannot be "
                                          "restarted: try larger `krylovdim`")
                        break
            restore_arnoldi_form(U, H, f, keep)
            for j in range(1, K + 1):                # copy!(rayleighquotient(fact), H): packed Hessenberg
                for i in range(1, min(j + 1, K) + 1):
                    fact.H[hidx(i, j)] = float(H[i - 1, j - 1])
            B = fact.basis()
            basistransform_(B, U[:, :keep])
            B[keep] = B[keep].scale_(1 / beta, fact.residual())
            fact = ar.shrink_(fact, keep)
            numiter += 1
    return T, U, fact, converged, numiter, numops


def _howmany_actual(T, fact, howmany, converged):
    hm = howmany
    if howmany < len(fact) and T[howmany, howmany - 1] != 0:
        hm += 1                                      # keep conjugate pairs together
    elif T.shape[0] < howmany:
        hm = T.shape[0]
    if converged > howmany:
        hm = converged
    return hm


def schursolve(A, x0: B200Vec, howmany: int, which: str, alg: Arnoldi):
    """schursolve(A, x₀, howmany, which, alg::Arnoldi) -> (T, vecs, vals, info) — arnoldi.jl:110-145."""
    T, U, fact, converged, numiter, numops = _schursolve(A, x0, howmany, which, alg)
    hm = _howmany_actual(T, fact, howmany, converged)
    TT = np.array(T[:hm, :hm])
    values = schur2eigvals(TT)
    B = fact.basis()
    vectors = [B * np.ascontiguousarray(U[:, i]) for i in range(hm)]
    r = fact.residual()
    residuals = [r.scale(float(U[-1, i])) for i in range(hm)]
    normres = np.array([fact.normres() * abs(U[-1, i]) for i in range(hm)])
    _warn(alg, "schursolve", converged, howmany, numiter, normres, numops)
    return TT, vectors, values, ConvergenceInfo(converged, residuals, normres, numiter, numops)


def eigsolve_arnoldi(A, x0: B200Vec, howmany: int, which: str, alg: Arnoldi, to_host: bool = False):
    """eigsolve(A, x₀, howmany, which, alg::Arnoldi) — arnoldi.jl:147-184.  `to_host` streams every Ritz
    vector / residual out as a complex numpy array as soon as it is formed (host-buffer entry: the number
    of converged pairs, hence of returned vectors, is not known when the slab is sized)."""
    T, U, fact, converged, numiter, numops = _schursolve(A, x0, howmany, which, alg)
    hm = _howmany_actual(T, fact, howmany, converged)
    TT = np.array(T[:hm, :hm])
    values = schur2eigvals(TT)
    V = U[:, :hm] @ schur2eigvecs(TT)
    B = fact.basis()
    r = fact.residual()
    vectors, residuals = [], []
    for i in range(hm):
        vre, vim = np.ascontiguousarray(V[:, i].real), np.ascontiguousarray(V[:, i].imag)
        vec = ComplexVec(B * vre, B * vim)
        vectors.append(vec.to_host() if to_host else vec)
        del vec
        res = ComplexVec(r.scale(float(vre[-1])), r.scale(float(vim[-1])))
        residuals.append(res.to_host() if to_host else res)
        del res
    normres = np.array([fact.normres() * abs(V[-1, i]) for i in range(hm)])
    _warn(alg, "eigsolve", converged, howmany, numiter, normres, numops)
    return values, vectors, ConvergenceInfo(converged, residuals, normres, numiter, numops)


def realeigsolve(A, x0: B200Vec, howmany: int, which: str, alg: Arnoldi):
    """realeigsolve(A, x₀, howmany, which, alg::Arnoldi) — arnoldi.jl:293-349: for operators whose
    spectrum is known to be real.  2×2 Schur blocks among the requested values are flattened with a
    warning (the imaginary parts are dropped); eigenvalues and eigenvectors come back real."""
    import math
    T, U, fact, converged, numiter, numops = _schursolve(A, x0, howmany, which, alg)
    T = np.array(T)
    K = len(fact)
    i = 0
    while i < howmany:
        i += 1
        if i < K:
            if abs(T[i, i - 1]) > alg.tol and alg.verbosity >= WARN_LEVEL:
                impart = math.sqrt(max(-T[i, i - 1] * T[i - 1, i], 0.0))
                warnings.warn(f"2 x 2 Schur block at position {i} and {i + 1} detected, complex eigenvalues with "
                              f"imaginary part {impart} will be ignored by setting T[i+1,i] = {T[i, i - 1]} to zero.")
            T[i, i - 1] = 0
    while i < converged:
        i += 1
        if i < K:
            if abs(T[i, i - 1]) <= alg.tol:
                T[i, i - 1] = 0
            else:
                i -= 1
                break
    hm = min(i, T.shape[0])
    converged = min(converged, hm)
    TT = T[:hm, :hm]
    values = np.diag(TT).copy()
    V = U[:, :hm] @ schur2realeigvecs(TT)
    B = fact.basis()
    r = fact.residual()
    vectors = [B * np.ascontiguousarray(V[:, j]) for j in range(hm)]
    residuals = [r.scale(float(V[-1, j])) for j in range(hm)]
    normres = np.array([fact.normres() * abs(V[-1, j]) for j in range(hm)])
    _warn(alg, "realeigsolve", converged, hm, numiter, normres, numops)
    return values, vectors, ConvergenceInfo(converged, residuals, normres, numiter, numops)


def _warn(alg, name, converged, howmany, numiter, normres, numops):
    if converged < howmany and alg.verbosity >= WARN_LEVEL:
        warnings.warn(f"Arnoldi {name} stopped without convergence after {numiter} iterations: "
                      f"{converged} values converged, normres = {normres}, numops = {numops}")
