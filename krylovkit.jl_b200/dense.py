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


class EigSorter:
    """EigSorter(by; rev = false) — src/eigsolve/eigsolve.jl:181-192: eigenvalues λ that come first (last
    if `rev`) when sorted by `by(λ)`.  `by` maps a numpy array of eigenvalues to real keys."""

    def __init__(self, by, rev: bool = False):
        self.by, self.rev = by, bool(rev)

    def permutation(self, vals):
        keys = np.asarray(self.by(np.asarray(vals)), dtype=np.float64)
        return np.argsort(-keys if self.rev else keys, kind="stable")      # stable like sortperm

    def __repr__(self):
        return f"EigSorter({getattr(self.by, '__name__', self.by)!r}, rev={self.rev})"


def eigsort(which):
    """eigsort — src/eigsolve/eigsolve.jl:335-355 (real spectra)."""
    if isinstance(which, EigSorter):
        return which.permutation
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


# ---- real Schur machinery for the Arnoldi drivers — dense/linalg.jl:150-393 ----------------------

def schur2eigvals(T: np.ndarray) -> np.ndarray:
    """schur2eigvals(T::Real) — dense/linalg.jl:166-189.  The first row of a 2×2 block carries +Im."""
    n = T.shape[0]
    D = np.zeros(n, dtype=np.complex128)
    i = 0
    while i < n:
        if i < n - 1 and T[i + 1, i] != 0:
            halftr = (T[i, i] + T[i + 1, i + 1]) / 2
            diff = (T[i, i] - T[i + 1, i + 1]) / 2
            im = math.sqrt(-(diff * diff + T[i, i + 1] * T[i + 1, i]))
            D[i], D[i + 1] = complex(halftr, im), complex(halftr, -im)
            i += 2
        else:
            D[i] = T[i, i]
            i += 1
    return D


def hschur(H: np.ndarray):
    """hschur!(H, Z) — dense/linalg.jl:152-154: real Schur form H = Z T Zᵀ (LAPACK, through scipy)."""
    from scipy.linalg import schur
    T, Z = schur(np.asarray(H, dtype=np.float64), output="real")
    return T, Z, schur2eigvals(T)


def eigsort_complex(which: str):
    """eigsort for complex Ritz values — eigsolve/eigsolve.jl:335-355; stable like sortperm."""
    if isinstance(which, EigSorter):
        return which.permutation
    keys = {"LM": lambda v: -np.abs(v), "LR": lambda v: -v.real, "SR": lambda v: v.real,
            "LI": lambda v: -v.imag, "SI": lambda v: v.imag}
    if which not in keys:
        raise ValueError(f"invalid specification of which eigenvalues to target: which = {which}")
    return lambda vals: np.argsort(keys[which](np.asarray(vals)), kind="stable")


def permuteschur(T: np.ndarray, Q: np.ndarray, order):
    """permuteschur!(T, Q, order) — dense/linalg.jl:356-386: reorder the diagonal blocks of a real
    Schur form with LAPACK trexc; a 2×2 block moves as a unit and may not be split."""
    T = np.asfortranarray(T, dtype=np.float64)
    Q = np.asfortranarray(Q, dtype=np.float64)
    n = T.shape[0]
    p = [int(v) + 1 for v in order]                      # trexc counts from one
    i = 0
    while i < len(p):
        ifirst = p[i]
        width = 1 if (ifirst == n or T[ifirst, ifirst - 1] == 0) else 2
        if width == 2 and (i + 1 >= len(p) or p[i + 1] != ifirst + 1):
            raise ValueError("cannot split 2x2 blocks when permuting schur decomposition")
        T, Q, info = lapack.dtrexc(T, Q, ifirst, i + 1)
        if info != 0:
            raise RuntimeError(f"dtrexc failed: info = {info}")
        for k in range(i + width, len(p)):
            if p[k] < p[i]:
                p[k] += width
        i += width
    return T, Q, schur2eigvals(T)


def schur2eigvecs(T: np.ndarray) -> np.ndarray:
    """schur2eigvecs(T::Real) — dense/linalg.jl:223-246: unit-norm complex eigenvectors of a real
    quasi-triangular T, column i belonging to schur2eigvals(T)[i].  The reference calls LAPACK trevc;
    here T is rotated to complex triangular form (rsf2csf) and back-substituted."""
    from scipy.linalg import rsf2csf
    n = T.shape[0]
    Tc, Zc = rsf2csf(np.asarray(T, dtype=np.float64), np.eye(n))
    smin = np.finfo(np.float64).eps * max(float(np.abs(Tc).max()), 1.0)
    X = np.zeros((n, n), dtype=np.complex128)
    for k in range(n):
        lam = Tc[k, k]
        X[k, k] = 1.0
        for i in range(k - 1, -1, -1):
            den = Tc[i, i] - lam
            if abs(den) < smin:
                den = smin
            X[i, k] = -(Tc[i, i + 1:k + 1] @ X[i + 1:k + 1, k]) / den
    W = Zc @ X
    W /= np.linalg.norm(W, axis=0)
    V = np.empty_like(W)
    i = 0
    while i < n:
        if i < n - 1 and T[i + 1, i] != 0:
            k = i if Tc[i, i].imag > 0 else i + 1          # whichever slot holds the +Im eigenvalue
            V[:, i] = W[:, k]
            V[:, i + 1] = np.conj(W[:, k])
            i += 2
        else:
            V[:, i] = W[:, i]
            i += 1
    return V


def restore_arnoldi_form(U: np.ndarray, H: np.ndarray, f: np.ndarray, keep: int):
    """_restorearnoldiform!(U, H, f, keep) — eigsolve/arnoldi.jl:468-481: Householder sweep that turns
    the truncated Krylov-Schur relation back into Arnoldi (Hessenberg) form."""
    H[keep, :keep] = f[:keep]
    for j in range(keep - 1, -1, -1):
        h, nu = householder_row(H, j + 1, range(0, j + 1), j)
        H[j + 1, j] = nu
        H[j + 1, :j] = 0
        lmul_householder(h, H)
        rmul_householder(H, h, slice(0, j + 1))
        rmul_householder(U, h)


def schur2realeigvecs(T: np.ndarray) -> np.ndarray:
    """schur2realeigvecs — dense/linalg.jl:247-257: real eigenvectors of an upper triangular T
    (LAPACK trevc in the reference; scipy's triangular solver here), unit 2-norm columns."""
    from scipy.linalg import solve_triangular
    T = np.asarray(T, dtype=np.float64)
    n = T.shape[0]
    if np.any(np.diag(T, -1) != 0):
        raise ValueError("T must be upper triangular")
    smin = np.finfo(np.float64).eps * max(float(np.abs(T).max()), 1.0)
    V = np.zeros((n, n))
    for k in range(n):
        V[k, k] = 1.0
        if k > 0:
            M = T[:k, :k] - T[k, k] * np.eye(k)
            d = np.diag(M).copy()
            d[np.abs(d) < smin] = smin                  # trevc-style perturbation of (near-)repeated values
            M[np.diag_indices(k)] = d
            V[:k, k] = solve_triangular(M, -T[:k, k], lower=False)
        V[:, k] /= np.linalg.norm(V[:, k])
    return V
