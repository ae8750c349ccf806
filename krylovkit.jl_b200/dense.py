"""Small dense k x k host algebra — mirror of src/dense/{linalg,reflector,givens,packedhessenberg}.jl.

These stay on the host in the reference too (LAPACK through ccall): O(k³)
work on the projected problem.  Only the *basis-side* applications (rmul! on an
OrthonormalBasis) go to the device, through orthonormal.py.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.linalg import lapack


class Householder:
    """Elementary reflector H = I - β v vᴴ acting on index range r (0-based list) —
    dense/reflector.jl:2-8."""

    __slots__ = ("beta", "v", "r")

    def __init__(self, beta, v, r):
        self.beta, self.v, self.r = beta, v, list(r)


def _householder(x, i):
    """_householder!(v, i) — dense/reflector.jl:34-65 (real scalars): reflect x onto e_i with
    a positive entry ν = ‖x‖."""
    v = np.array(x, dtype=np.float64)
    sigma = float(np.sum(v[:i] ** 2) + np.sum(v[i + 1:] ** 2))
    vi = float(v[i])
    nu = math.sqrt(vi * vi + sigma)
    if sigma == 0.0 and vi == nu:
        return 0.0, v, nu
    if vi < 0:
        vi = vi - nu
    else:
        vi = -sigma / (vi + nu)
    v[:i] /= vi
    v[i + 1:] /= vi
    v[i] = 1.0
    return -vi / nu, v, nu


def householder_row(A, row, r, k):
    """householder(A, row, r, k): zeros A[row, r] except A[row, k] upon rmul!(A, h') —
    dense/reflector.jl:24-29."""
    r = list(r)
    beta, v, nu = _householder(A[row, r], r.index(k))
    return Householder(beta, v, r), nu


def householder_col(A, r, col, k):
    """householder(A, r, col, k): zeros A[r, col] except A[k, col] upon lmul!(h, A) —
    dense/reflector.jl:17-22."""
    r = list(r)
    beta, v, nu = _householder(A[r, col], r.index(k))
    return Householder(beta, v, r), nu


def lmul_householder(H: Householder, A, cols=None):
    """lmul!(H, A[, cols]) — dense/reflector.jl:85-106."""
    if H.beta == 0:
        return A
    cols = slice(None) if cols is None else (slice(cols.start, cols.stop) if isinstance(cols, range) else cols)
    r = H.r
    if r == list(range(r[0], r[0] + len(r))):      # contiguous index range -> views, no gather
        sub = A[r[0]:r[0] + len(r), cols]
        mu = H.beta * (H.v @ sub)                  # one μ per column (reflector.jl:93-99)
        sub -= np.outer(H.v, mu)
    else:
        sub = A[np.ix_(r, np.arange(A.shape[1])[cols])]
        mu = H.beta * (H.v @ sub)
        A[np.ix_(r, np.arange(A.shape[1])[cols])] = sub - np.outer(H.v, mu)
    return A


def rmul_householder(A, H: Householder, rows=None):
    """rmul!(A, H[, rows]) — dense/reflector.jl:107-142 (real: H' == H)."""
    if H.beta == 0:
        return A
    rows = slice(None) if rows is None else rows
    sub = A[rows, :]
    w = sub[:, H.r] @ H.v
    sub[:, H.r] -= np.outer(w, H.beta * H.v)
    return A


def givens(f: float, g: float):
    """LinearAlgebra.givens(f, g, i1, i2): (c, s, r) with [c s; -s c]·[f; g] = [r; 0]."""
    c, s, r = lapack.dlartg(f, g)
    return float(c), float(s), float(r)


def tridiageigh(dv, ev):
    """tridiageigh!(SymTridiagonal(dv, ev), Z) -> stegr! — dense/linalg.jl:109-115, 396-458.
    LAPACK dstemr (= dstegr's MRRR algorithm) through scipy."""
    n = len(dv)
    if n == 1:
        return np.array(dv, dtype=np.float64), np.ones((1, 1))
    e = np.zeros(n)
    e[: n - 1] = ev[: n - 1]
    m, w, z, info = lapack.dstemr(np.array(dv, dtype=np.float64), e, 0, 0.0, 0.0, 1, n, compute_v=1)
    if info != 0:
        raise RuntimeError(f"dstemr failed: info = {info}")
    return w[:n].copy(), z[:, :n].copy()


def bidiagsvd_lower(alphas, betas):
    """bidiagsvd!(Bidiagonal(αs, βs, :L), P, Q) -> (P, S, Q), B = P Diag(S) Q — dense/linalg.jl:123-130
    (LAPACK bdsqr in the reference; here the SVD of the k x k bidiagonal: same values,
    vectors up to sign)."""
    K = len(alphas)
    B = np.diag(np.asarray(alphas, dtype=np.float64))
    for i in range(K - 1):
        B[i + 1, i] = betas[i]
    return np.linalg.svd(B)


def eigsort(which: str):
    """eigsort — src/eigsolve/eigsolve.jl:335-355 (real spectra)."""
    if which == "SR":
        return lambda d: np.argsort(d, kind="stable")
    if which == "LR":
        return lambda d: np.argsort(-d, kind="stable")
    if which == "LM":
        return lambda d: np.argsort(-np.abs(d), kind="stable")
    raise ValueError(f"invalid specification of which eigenvalues to target: which = {which}")


def permuteeig(D, V, p):
    """permuteeig!(D, V, perm) — dense/linalg.jl:306-333."""
    return D[p].copy(), V[:, p].copy()


def ldiv_upper(R, y, k):
    """ldiv!(UpperTriangular(R), y, 1:k) — dense/linalg.jl:96-106."""
    for j in range(k - 1, -1, -1):
        if R[j, j] == 0:
            raise ZeroDivisionError(f"SingularException({j + 1})")
        yj = y[j] / R[j, j]
        y[j] = yj
        y[:j] -= R[:j, j] * yj
    return y


def hidx(i: int, j: int) -> int:
    """PackedHessenberg (1-based i <= j+1) -> 0-based index into data — packedhessenberg.jl:32-39."""
    return ((j * j + j - 2) >> 1) + i - 1
