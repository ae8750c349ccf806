"""CPU oracle — a plain numpy/scipy restatement of KrylovKit.jl's Krylov hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under krylovkit.jl_b200/ may import this module; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do,
and only as the checker (or the timed CPU baseline), never as the product.

What it restates (operation order follows the cited reference lines; paths relative to
the KrylovKit.jl v0.10.4 tree):
  * VectorInterface calls on plain arrays: inner = dot, norm = nrm2, add!! = axpby, scale!!
  * src/orthonormal.jl:88-150 project!!/unproject!!, :210-229 rank1update!, :291-321
    basistransform!, :372-489 the six orthogonalize!! variants, :522-527 orthonormalize!!
  * src/dense/reflector.jl:31-65 Householder, :67-154 lmul!/rmul!, dense/givens.jl:12-36
  * src/factorizations/lanczos.jl:180-376, arnoldi.jl:135-245, gkl.jl:183-404
  * src/factorizations/blocklanczos.jl:43-52, 277-284, 312-353 block primitives
  * src/eigsolve/lanczos.jl:1-155, src/linsolve/gmres.jl:1-151, src/eigsolve/svdsolve.jl:144-314
  * src/dense/linalg.jl:96-106 ldiv!, :306-333 permuteeig!, packedhessenberg.jl:32-48
  * widened rows: src/linsolve/cg.jl, bicgstab.jl, src/lssolve/lsmr.jl, src/factorizations/blocklanczos.jl
    + src/eigsolve/blocklanczos.jl, src/eigsolve/arnoldi.jl (schursolve / eigsolve / realeigsolve) with
    src/dense/linalg.jl:150-393 (real Schur helpers), src/matrixfun/expintegrator.jl:101-323
  The n-length loops also exist in C/OpenMP (oracle/csrc/kernels.c, plugged in by oracle/native.py) for
  the timed CPU baseline; they are checked against the numpy primitives here.

Arithmetic that lives outside the reference tree (VectorInterface.jl 0.5/0.6, Julia's
LinearAlgebra/OpenBLAS, SparseArrays — Project.toml:31-48, no Manifest => versions
unpinned) is restated by numpy/scipy calls with the same mathematical definition:
`A @ x` (scipy CSR, single-threaded like SparseArrays' CSC product), np.dot, np.linalg.norm,
LAPACK dstemr through scipy (the reference calls dstegr, which is dstemr with the same
MRRR algorithm, dense/linalg.jl:396-458), numpy SVD of the k x k bidiagonal instead of
LAPACK dbdsqr (values identical, vectors up to sign).

PINNING.  The reference cannot run here (no Julia in the image).  The oracle is pinned
against the reference's own known-answer fixtures (tests/test_oracle_golden.py):
the literal 71x71 matrix of issue #143 (test/issues.jl:39-129), eigsolve([1 0;0 1])
(issues.jl:32-36), the toric-code ground energy -16 (test/eigsolve.jl:471-549) and the
invariants the reference asserts after every expand!/shrink! (test/factorize.jl:140-158,
185-203, 285-309), lssolve(I(2), [1, 0]) of issue #133 (issues.jl:21-29), and — for every widened
solver — the assertions of the reference's own test file for it (test/linsolve.jl, lssolve.jl,
eigsolve.jl, expintegrator.jl), run on the oracle.  Independent implementations agree too: SciPy's LSMR
and BiCGStab iterate for iterate, expm_multiply, ARPACK.  Bit-level parity with the Julia/OpenBLAS
reduction order is "parity unpinned": no reference test fixes it.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp
from scipy.linalg import blas as _blas
from scipy.linalg import lapack

# orthogonalizer tags (src/algorithms.jl:17-80)
CGS, MGS, CGS2, MGS2, CGSIR, MGSIR = range(6)
ETA_DEFAULT = 1.0 / math.sqrt(2.0)


@dataclass(frozen=True)
class Orth:
    tag: int
    eta: float = ETA_DEFAULT


# ------------------------------------------------------------------ inputs --------------

def splitmix64(z):
    """counter RNG shared with the device (csrc/common.cuh splitmix64)."""
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def splitmix_vector(seed: int, n: int, offset: int = 0, dtype=np.float64):
    i = np.arange(n, dtype=np.uint64) + np.uint64(offset) + np.uint64(seed)
    return ((splitmix64(i) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)).astype(dtype)


def stencil_matrix(nx, ny, nz=1, coeffs=(4.0, -1.0, -1.0, -1.0, -1.0, -1.0, -1.0), dtype=np.float64):
    """Dirichlet stencil on an nx*ny*nz grid (x fastest): (centre, west, east, south, north, down, up)."""
    c, w, e, s, n_, d, u = coeffs
    Ix, Iy, Iz = sp.identity(nx), sp.identity(ny), sp.identity(nz)
    ex = sp.diags([w * np.ones(max(nx - 1, 0)), e * np.ones(max(nx - 1, 0))], [-1, 1], shape=(nx, nx))
    ey = sp.diags([s * np.ones(max(ny - 1, 0)), n_ * np.ones(max(ny - 1, 0))], [-1, 1], shape=(ny, ny))
    A = sp.kron(Iz, sp.kron(Iy, ex)) + sp.kron(Iz, sp.kron(ey, Ix)) + c * sp.identity(nx * ny * nz)
    if nz > 1:
        ez = sp.diags([d * np.ones(nz - 1), u * np.ones(nz - 1)], [-1, 1], shape=(nz, nz))
        A = A + sp.kron(ez, sp.kron(Iy, Ix))
    A = A.tocsr().astype(dtype)
    A.sort_indices()
    return A


def laplace_eigenvalues(nx, ny, nz=1):
    """closed-form spectrum of the Dirichlet 5-/7-point Laplacian."""
    lx = 2 - 2 * np.cos(np.arange(1, nx + 1) * np.pi / (nx + 1))
    ly = 2 - 2 * np.cos(np.arange(1, ny + 1) * np.pi / (ny + 1))
    lam = lx[:, None] + ly[None, :]
    if nz > 1:
        lz = 2 - 2 * np.cos(np.arange(1, nz + 1) * np.pi / (nz + 1))
        lam = lam[:, :, None] + lz[None, None, :]
    return np.sort(lam.ravel())


def dense_splitmix(seed, m, n, dtype=np.float32, row0=0, m_global=None):
    mg = m if m_global is None else m_global
    i = np.arange(m, dtype=np.uint64)[:, None] + np.uint64(row0)
    j = np.arange(n, dtype=np.uint64)[None, :] * np.uint64(mg)
    z = splitmix64(i + j + np.uint64(seed))
    return np.asfortranarray(((z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) - 0.5)
                             .astype(dtype))


# ------------------------------------------------------------------ apply (src/apply.jl) --

def apply(op, x, a0=0.0, a1=1.0):
    """apply.jl:1-11."""
    y = op @ x if not callable(op) else op(x)
    if a0 != 0 or a1 != 1:
        y = a1 * y + a0 * x          # add!!(y, x, α₀, α₁): y = α₁*y + α₀*x
    return y


def apply_normal(op, x):
    return op[0](x) if isinstance(op, tuple) else (op(x, False) if callable(op) else op @ x)


def apply_adjoint(op, x):
    return op[1](x) if isinstance(op, tuple) else (op(x, True) if callable(op) else op.T @ x)


def inner(x, y):
    return float(np.dot(x, y))


def _axpy(y, a, x):
    """y <- y + a*x IN PLACE through BLAS axpy (what VectorInterface.add!! lowers to for Arrays:
    LinearAlgebra.axpy!).  y must be an array this module owns (fresh result of apply / copy)."""
    if y.dtype == np.float64 and x.dtype == np.float64 and y.flags.c_contiguous and x.flags.c_contiguous:
        return _blas.daxpy(x, y, a=float(a))
    if y.dtype == np.float32 and x.dtype == np.float32 and y.flags.c_contiguous and x.flags.c_contiguous:
        return _blas.saxpy(x, y, a=float(a))
    y += y.dtype.type(a) * x
    return y


def norm(x):
    return float(np.linalg.norm(x))


# ------------------------------------------------------------------ orthonormal.jl --------

def project(y, b, x, alpha=1.0, beta=0.0, r=None):
    """orthonormal.jl:88-118."""
    r = range(len(b)) if r is None else r
    for j, rj in enumerate(r):
        if beta == 0:
            y[j] = alpha * inner(b[rj], x)
        else:
            y[j] = beta * y[j] + alpha * inner(b[rj], x)
    return y


def scale_(x, a):
    """scale!!(x, a) IN PLACE — for arrays the caller owns (lanczos.jl:257 `scale!!(r, 1/β)`)."""
    x *= x.dtype.type(a)
    return x


def unproject(y, b, x, alpha=1.0, beta=0.0, r=None, inplace=False):
    """orthonormal.jl:132-150 (generic BLAS-1 path).  `inplace` (beta == 1 only): update y itself, as the
    reference's unproject!! does — for callers that own y; by default the input is left untouched."""
    r = range(len(b)) if r is None else r
    if beta == 0:
        y = np.zeros_like(y)         # hard zero
    elif beta != 1:
        y = y * beta
    elif not inplace:
        y = y.copy()                 # the oracle's callers keep their input; one copy, then in-place axpys
    for i, ri in enumerate(r):
        y = _axpy(y, alpha * x[i], b[ri])
    return y


def rank1update(b, y, x, alpha=1.0, beta=1.0, r=None):
    """orthonormal.jl:210-229."""
    r = range(len(b)) if r is None else r
    for i, ri in enumerate(r):
        if beta == 1:
            b[ri] = b[ri] + (alpha * np.conj(x[i])) * y
        elif beta == 0:
            b[ri] = (alpha * np.conj(x[i])) * y
        else:
            b[ri] = beta * b[ri] + (alpha * np.conj(x[i])) * y
    return b


def basistransform(b, U):
    """orthonormal.jl:291-321."""
    m, n = U.shape
    assert m == len(b)
    b2 = []
    for j in range(n):
        v = b[0] * U[0, j]
        for i in range(1, m):
            v = v + b[i] * U[i, j]
        b2.append(v)
    for j in range(n):
        b[j] = b2[j]
    return b


def _cgs_pass(v, b, x, own=False):
    x = project(x, b, v)             # orthonormal.jl:381
    v = unproject(v, b, x, -1, 1, inplace=own)    # :382
    return v, x


def _mgs_pass(v, b, x, accumulate):
    v = v.copy()
    for i, q in enumerate(b):        # orthonormal.jl:417-421 / 427-431
        s = inner(q, v)
        v = _axpy(v, -s, q)
        if accumulate:
            x[i] += s
        else:
            x[i] = s
    return v, x


def orthogonalize(v, b, x, orth: Orth):
    """orthogonalize!!(v, b, x, alg) — orthonormal.jl:378-452."""
    t = orth.tag
    if t == CGS:
        return _cgs_pass(v, b, x)
    if t == CGS2:
        v, x = _cgs_pass(v, b, x)
        s = np.empty_like(x)
        v, s = _cgs_pass(v, b, s)    # reorthogonalize!! :385-393
        x += s
        return v, x
    if t == CGSIR:
        nold = norm(v)
        v, x = _cgs_pass(v, b, x)
        nnew = norm(v)
        while np.finfo(np.float64).eps < nnew < orth.eta * nold:
            nold = nnew
            s = np.empty_like(x)
            v, s = _cgs_pass(v, b, s)
            x += s
            nnew = norm(v)
        return v, x
    if t == MGS:
        return _mgs_pass(v, b, x, False)
    if t == MGS2:
        v, x = _mgs_pass(v, b, x, False)
        return _mgs_pass(v, b, x, True)
    if t == MGSIR:
        nold = norm(v)
        v, x = _mgs_pass(v, b, x, False)
        nnew = norm(v)
        while np.finfo(np.float64).eps < nnew < orth.eta * nold:
            nold = nnew
            v, x = _mgs_pass(v, b, x, True)
            nnew = norm(v)
        return v, x
    raise ValueError(t)


def orthogonalize_vec(v, q, orth: Orth, eps=np.finfo(np.float64).eps):
    """orthogonalize!!(v, q, alg) against one vector — orthonormal.jl:455-489."""
    t = orth.tag
    if t in (CGS, MGS):
        s = inner(q, v)
        return v - s * q, s
    if t in (CGS2, MGS2):
        s = inner(q, v)
        v = v - s * q
        ds = inner(q, v)
        v = v - ds * q
        return v, s + ds
    nold = norm(v)
    s = inner(q, v)
    v = v - s * q
    nnew = norm(v)
    while eps < nnew < orth.eta * nold:
        nold = nnew
        ds = inner(q, v)
        v = v - ds * q
        s += ds
        nnew = norm(v)
    return v, s


def orthonormalize(v, b, x, orth):
    """orthonormal.jl:522-527."""
    v, x = orthogonalize(v, b, x, orth)
    beta = norm(v)
    return v / beta, beta, x


# ------------------------------------------------------------------ dense helpers ---------

def householder_vec(x, i):
    """_householder!(v, i) — dense/reflector.jl:34-65 (real).  Returns (β, v, ν)."""
    v = np.array(x, dtype=np.float64)
    sigma = float(np.sum(v[:i] ** 2) + np.sum(v[i + 1:] ** 2))
    vi = v[i]
    nu = math.sqrt(vi * vi + sigma)
    if sigma == 0 and vi == nu:
        return 0.0, v, nu
    if vi < 0:
        vi = vi - nu
    else:
        vi = (-sigma) / (vi + nu)    # ((vi - conj(vi))ν - σ)/(conj(vi)+ν) for real vi
    v[:i] /= vi
    v[i + 1:] /= vi
    v[i] = 1.0
    beta = -vi / nu
    return beta, v, nu


def householder_lmul(beta, v, r, A):
    """lmul!(H, A) — reflector.jl:85-106: rows r of A."""
    if beta == 0:
        return A
    r = list(r)
    for k in range(A.shape[1]):
        mu = beta * float(np.dot(v, A[r, k]))
        A[r, k] -= mu * v
    return A


def householder_rmul(A, beta, v, r):
    """rmul!(A, H) — reflector.jl:107-142: columns r of A (H already adjointed if needed; real)."""
    if beta == 0:
        return A
    r = list(r)
    w = A[:, r] @ v
    A[:, r] -= np.outer(w, beta * v)
    return A


def householder_rmul_basis(b, beta, v, r):
    """rmul!(b::OrthonormalBasis, H) — reflector.jl:143-154."""
    if beta == 0:
        return b
    r = list(r)
    w = unproject(np.zeros_like(b[r[0]]), b, v, 1, 0, r)
    return rank1update(b, w, v, -beta, 1, r)


def givens(f, g):
    """LinearAlgebra.givens(f, g, i1, i2) -> (c, s, r) with [c s; -s c][f; g] = [r; 0] (dlartg)."""
    c, s, r = lapack.dlartg(f, g)
    return float(c), float(s), float(r)


def givens_rmul_basis(b, i1, i2, c, s):
    """_rmul!(b, G) — dense/givens.jl:31-36."""
    q1, q2 = b[i1], b[i2]
    b[i1], b[i2] = c * q1 - s * q2, s * q1 + c * q2
    return b


def tridiageigh(dv, ev):
    """tridiageigh! -> stegr! (dense/linalg.jl:109-115, 396-458); dstemr here."""
    n = len(dv)
    if n == 1:
        return np.array(dv, dtype=np.float64), np.ones((1, 1))
    e = np.zeros(n)
    e[: n - 1] = ev[: n - 1]
    m, w, z, info = lapack.dstemr(np.array(dv, dtype=np.float64), e, 0, 0.0, 0.0, 1, n, compute_v=1)
    assert info == 0
    return w[:n].copy(), z[:, :n].copy()


def permuteeig(D, V, p):
    """permuteeig! — dense/linalg.jl:306-333 (net effect: D[p], V[:, p])."""
    return D[p].copy(), V[:, p].copy()


def eigsort(which):
    """eigsolve.jl:335-355."""
    if which == "SR":
        return lambda d: np.argsort(d, kind="stable")
    if which == "LR":
        return lambda d: np.argsort(-d, kind="stable")
    if which == "LM":
        return lambda d: np.argsort(-np.abs(d), kind="stable")
    raise ValueError(which)


def ldiv_upper(R, y, k):
    """ldiv!(UpperTriangular(R), y, 1:k) — dense/linalg.jl:96-106."""
    for j in range(k - 1, -1, -1):
        if R[j, j] == 0:
            raise ZeroDivisionError("singular")
        yj = y[j] / R[j, j]
        y[j] = yj
        y[:j] -= R[:j, j] * yj
    return y


# ------------------------------------------------------------------ Lanczos ---------------

@dataclass
class LanczosFact:
    k: int
    V: list
    alphas: list
    betas: list
    r: np.ndarray

    def normres(self):
        return self.betas[-1]


def lanczos_initialize(A, x0, orth: Orth):
    """initialize(iter::LanczosIterator) — factorizations/lanczos.jl:180-222."""
    beta0 = norm(x0)
    if beta0 == 0:
        raise ValueError("initial vector should not have norm zero")
    Ax0 = apply(A, x0)
    alpha = inner(x0, Ax0) / (beta0 * beta0)
    v = x0 * (1 / beta0)             # add!!(scale(Ax₀, 0), x₀, 1/β₀)
    r = Ax0 * (1 / beta0)
    betaold = norm(r)
    r = r - alpha * v
    beta = norm(r)
    if orth.tag in (CGS2, MGS2):
        dalpha = inner(v, r)
        alpha += dalpha
        r = r - dalpha * v
        beta = norm(r)
    elif orth.tag in (CGSIR, MGSIR):
        while np.finfo(np.float64).eps < beta < orth.eta * betaold:
            betaold = beta
            dalpha = inner(v, r)
            alpha += dalpha
            r = r - dalpha * v
            beta = norm(r)
    return LanczosFact(1, [v], [alpha], [beta], r)


def lanczos_recurrence(A, V, beta, orth: Orth):
    """lanczosrecurrence — factorizations/lanczos.jl:295-376."""
    t = orth.tag
    eps = np.finfo(np.float64).eps
    v = V[-1]
    w = apply(A, v)
    if t == CGS:
        alpha = inner(v, w)
        w = w - beta * V[-2]
        w = w - alpha * v
        return w, alpha, norm(w)
    if t == MGS:
        w = w - beta * V[-2]
        alpha = inner(v, w)
        w = w - alpha * v
        return w, alpha, norm(w)
    if t == CGS2:
        alpha = inner(v, w)
        w = _axpy(w, -beta, V[-2])   # w is the fresh result of apply: mutate in place like add!!
        w = _axpy(w, -alpha, v)
        s = np.empty(len(V))
        w, s = _cgs_pass(w, V, s, own=True)
        alpha += s[-1]
        return w, alpha, norm(w)
    if t == MGS2:
        w = _axpy(w, -beta, V[-2])
        alpha = inner(v, w)
        w = _axpy(w, -alpha, v)
        s = alpha
        for q in V:
            s = inner(q, w)
            w = _axpy(w, -s, q)
        alpha += s
        return w, alpha, norm(w)
    if t == CGSIR:
        alpha = inner(v, w)
        w = w - beta * V[-2]
        w = w - alpha * v
        ab2 = alpha * alpha + beta * beta
        beta = norm(w)
        nold = math.sqrt(beta * beta + ab2)
        while eps < beta < orth.eta * nold:
            nold = beta
            s = np.empty(len(V))
            w, s = _cgs_pass(w, V, s)
            alpha += s[-1]
            beta = norm(w)
        return w, alpha, beta
    if t == MGSIR:
        w = w - beta * V[-2]
        alpha = inner(v, w)
        w = w - alpha * v
        ab2 = alpha * alpha + beta * beta
        beta = norm(w)
        nold = math.sqrt(beta * beta + ab2)
        while eps < beta < orth.eta * nold:
            nold = beta
            s = 0.0
            for q in V:
                s = inner(q, w)
                w = w - s * q
            alpha += s
            beta = norm(w)
        return w, alpha, beta
    raise ValueError(t)


def lanczos_expand(A, f: LanczosFact, orth: Orth):
    """expand! — factorizations/lanczos.jl:250-272."""
    betaold = f.normres()
    f.V.append(scale_(f.r, 1 / betaold))       # the residual's storage becomes the new basis vector (:257)
    r, alpha, beta = lanczos_recurrence(A, f.V, betaold, orth)
    f.alphas.append(alpha)
    f.betas.append(beta)
    f.k += 1
    f.r = r
    return f


def lanczos_shrink(f: LanczosFact, k):
    """shrink! — factorizations/lanczos.jl:273-291."""
    if f.k <= k:
        return f
    while len(f.V) > k + 1:
        f.V.pop()
    r = f.V.pop()
    del f.alphas[k:]
    del f.betas[k:]
    f.k = k
    f.r = r * f.normres()
    return f


def eigsolve_lanczos(A, x0, howmany, which, krylovdim=30, maxiter=100, tol=1e-12,
                     orth: Orth = Orth(MGS2), eager=False, callback=None):
    """eigsolve(A, x₀, howmany, which, ::Lanczos) — src/eigsolve/lanczos.jl:1-155.
    Returns (values, vectors, info dict)."""
    if howmany > krylovdim:
        raise ValueError("krylov dimension too small")
    f = lanczos_initialize(A, x0, orth)
    numops, numiter = 1, 1
    HH = np.zeros((krylovdim + 1, krylovdim))
    converged = 0
    D = U = fvec = None
    while True:
        beta = f.normres()
        K = f.k
        if K == krylovdim or beta <= tol or (eager and K >= howmany):
            if K == 1:
                D = np.array([f.alphas[0]])
                U = np.ones((1, 1))
                fvec = np.array([beta])
                converged = int(beta <= tol)
            else:
                D, U = tridiageigh(f.alphas[:K], f.betas[:K - 1])
                p = eigsort(which)(D)
                D, U = permuteeig(D, U, p)
                fvec = U[K - 1, :] * beta
                converged = 0
                while converged < K and abs(fvec[converged]) <= tol:
                    converged += 1
            if converged >= howmany or beta <= tol:
                break
        if K < krylovdim:
            f = lanczos_expand(A, f, orth)
            numops += 1
            if callback:
                callback(f)
        else:
            if numiter == maxiter:
                break
            keep = (3 * krylovdim + 2 * converged) // 5
            H = HH[: keep + 1, :keep]
            H[:] = 0
            for j in range(keep):
                H[j, j] = D[j]
                H[keep, j] = fvec[j]
            for j in range(keep - 1, -1, -1):        # Julia j = keep:-1:1
                rr = list(range(0, j + 1))
                hb, hv, nu = householder_vec(H[j + 1, rr], j)
                H[j + 1, j] = nu
                H[j + 1, :j] = 0
                householder_lmul(hb, hv, rr, H)
                householder_rmul(H[: j + 1, :], hb, hv, rr)
                householder_rmul(U, hb, hv, rr)
            for j in range(keep):
                f.alphas[j] = H[j, j]
                f.betas[j] = H[j + 1, j]
            basistransform(f.V, U[:, :keep])
            f.V[keep] = f.r * (1 / beta)
            f = lanczos_shrink(f, keep)
            numiter += 1
    hm = howmany
    if converged > howmany:
        hm = converged
    elif len(D) < howmany:
        hm = len(D)
    values = D[:hm].copy()
    vectors = [unproject(np.zeros_like(f.V[0]), f.V, U[:, i]) for i in range(hm)]
    residuals = [f.r * U[-1, i] for i in range(hm)]
    normres = np.abs(fvec[:hm])
    return values, vectors, dict(converged=converged, residual=residuals, normres=normres,
                                 numiter=numiter, numops=numops, fact=f)


# ------------------------------------------------------------------ Arnoldi / GMRES -------

def hidx(i, j):
    """PackedHessenberg index (1-based i, j) -> 0-based data index — packedhessenberg.jl:32-39."""
    return ((j * j + j - 2) >> 1) + i - 1


@dataclass
class ArnoldiFact:
    k: int
    V: list
    H: list          # packed Hessenberg incl. trailing β
    r: np.ndarray

    def normres(self):
        return abs(self.H[-1])

    def h(self, i, j):
        return self.H[hidx(i, j)]


def arnoldi_initialize(A, x0, orth: Orth):
    """arnoldi.jl:135-175 (identical to the Lanczos initialisation)."""
    f = lanczos_initialize(A, x0, orth)
    return ArnoldiFact(1, f.V, [f.alphas[0], f.betas[0]], f.r)


def arnoldi_initialize_inplace(A, x0, f: ArnoldiFact, orth: Orth):
    """initialize! — arnoldi.jl:176-198."""
    while len(f.V) > 1:
        f.V.pop()
    f.V[0] = x0 * (1 / norm(x0))
    w = apply(A, f.V[0])
    r, alpha = orthogonalize_vec(w, f.V[0], orth)
    f.k = 1
    f.H = [alpha, norm(r)]
    f.r = r
    return f


def arnoldi_expand(A, f: ArnoldiFact, orth: Orth):
    """expand! — arnoldi.jl:199-219; arnoldirecurrence!! :239-245."""
    f.k += 1
    k = f.k
    beta = f.normres()
    f.V.append(f.r * (1 / beta))
    w = apply(A, f.V[-1])
    h = np.empty(k)
    r, h = orthogonalize(w, f.V, h, orth)
    f.H.extend(h.tolist())
    f.H.append(norm(r))
    f.r = r
    return f


def arnoldi_shrink(f: ArnoldiFact, k):
    """shrink! — factorizations/arnoldi.jl:220-236."""
    if f.k <= k:
        return f
    while len(f.V) > k + 1:
        f.V.pop()
    r = f.V.pop()
    del f.H[(k * k + 3 * k) >> 1:]
    f.k = k
    f.r = r * f.normres()
    return f


def arnoldi_rayleighquotient(f: ArnoldiFact):
    k = f.k
    Hm = np.zeros((k, k))
    for j in range(1, k + 1):
        for i in range(1, min(j + 1, k) + 1):
            Hm[i - 1, j - 1] = f.H[hidx(i, j)]
    return Hm


def schur2eigvals(T):
    """schur2eigvals(T::Real) — dense/linalg.jl:166-189: complex eigenvalues from the 1×1 / 2×2 diagonal
    blocks of a real quasi-triangular matrix; the first row of a 2×2 block carries +Im."""
    n = T.shape[0]
    D = np.zeros(n, dtype=np.complex128)
    for i in range(n):
        if i < n - 1 and T[i + 1, i] != 0:
            halftr = (T[i, i] + T[i + 1, i + 1]) / 2
            diff = (T[i, i] - T[i + 1, i + 1]) / 2
            d = diff * diff + T[i, i + 1] * T[i + 1, i]
            D[i] = halftr + 1j * math.sqrt(-d)
        elif i > 0 and T[i, i - 1] != 0:
            halftr = (T[i, i] + T[i - 1, i - 1]) / 2
            diff = -(T[i, i] - T[i - 1, i - 1]) / 2
            d = diff * diff + T[i, i - 1] * T[i - 1, i]
            D[i] = halftr - 1j * math.sqrt(-d)
        else:
            D[i] = T[i, i]
    return D


def hschur(H):
    """hschur!(H, Z) — dense/linalg.jl:152-154 (LAPACK hseqr): real Schur form H = Z T Zᵀ."""
    from scipy.linalg import schur
    T, Z = schur(H, output="real")
    return T, Z, schur2eigvals(T)


def eigsort_complex(which):
    """eigsort — eigsolve/eigsolve.jl:335-355 for complex values: stable permutation."""
    key = {"LM": lambda v: -np.abs(v), "LR": lambda v: -v.real, "SR": lambda v: v.real,
           "LI": lambda v: -v.imag, "SI": lambda v: v.imag}[which]
    return lambda vals: np.argsort(key(np.asarray(vals)), kind="stable")


def permuteschur(T, Q, order):
    """permuteschur!(T, Q, order) for real T — dense/linalg.jl:356-386: bring the diagonal blocks into
    the given order with LAPACK trexc, never splitting a 2×2 block."""
    T = np.asfortranarray(T, dtype=np.float64)
    Q = np.asfortranarray(Q, dtype=np.float64)
    n = T.shape[0]
    p = [int(v) + 1 for v in order]                      # 1-based like the reference
    i = 0
    while i < len(p):
        ifirst, ilast = p[i], i + 1
        single = ifirst == n or T[ifirst, ifirst - 1] == 0     # T[ifirst+1, ifirst] in 1-based terms
        if not single and (i + 1 >= len(p) or p[i + 1] != ifirst + 1):
            raise ValueError("cannot split 2x2 blocks when permuting schur decomposition")
        T, Q, info = lapack.dtrexc(T, Q, ifirst, ilast)
        if info != 0:
            raise RuntimeError(f"trexc failed: info = {info}")
        step = 1 if single else 2
        for k in range(i + step, len(p)):
            if p[k] < p[i]:
                p[k] += step
        i += step
    return T, Q, schur2eigvals(T)


def schur2eigvecs(T):
    """schur2eigvecs(T::Real) — dense/linalg.jl:223-246 (LAPACK trevc + normalisation): unit-norm complex
    eigenvectors of a real quasi-triangular matrix, column i ↔ schur2eigvals(T)[i]; the two columns of a
    2×2 block are complex conjugates.  Restated as a back substitution per diagonal block."""
    n = T.shape[0]
    vals = schur2eigvals(T)
    V = np.zeros((n, n), dtype=np.complex128)
    i = 0
    while i < n:
        two = i < n - 1 and T[i + 1, i] != 0
        lam = vals[i]
        x = np.zeros(n, dtype=np.complex128)
        if two:
            x[i], x[i + 1] = T[i, i + 1], lam - T[i, i]          # null vector of the 2×2 block minus λ
            top = i
            rhs = -(T[:top, i:i + 2] @ x[i:i + 2])
        else:
            x[i] = 1.0
            top = i
            rhs = -T[:top, i].astype(np.complex128)
        if top > 0:
            Lm = T[:top, :top] - lam * np.eye(top)
            smin = np.finfo(float).eps * max(np.abs(T).max(), 1.0)
            for d in range(top):                               # trevc-style perturbation of tiny pivots
                if abs(Lm[d, d]) < smin and not (d < top - 1 and T[d + 1, d] != 0) \
                        and not (d > 0 and T[d, d - 1] != 0):
                    Lm[d, d] = smin
            x[:top] = np.linalg.solve(Lm, rhs)
        x /= np.linalg.norm(x)
        V[:, i] = x
        if two:
            V[:, i + 1] = np.conj(x)
            i += 2
        else:
            i += 1
    return V


def restore_arnoldi_form(U, H, f, keep):
    """_restorearnoldiform!(U, H, f, keep) — eigsolve/arnoldi.jl:468-481."""
    for j in range(keep):
        H[keep, j] = f[j]
    for j in range(keep - 1, -1, -1):
        rr = list(range(j + 1))
        hb, hv, nu = householder_vec(H[j + 1, rr], j)
        H[j + 1, j] = nu
        H[j + 1, :j] = 0
        householder_lmul(hb, hv, rr, H)
        householder_rmul(H[: j + 1, :], hb, hv, rr)
        householder_rmul(U, hb, hv, rr)


def _schursolve(A, x0, howmany, which, krylovdim, maxiter, tol, orth, eager):
    """_schursolve — eigsolve/arnoldi.jl:351-452 (Krylov-Schur restarts in real arithmetic)."""
    if howmany > krylovdim:
        raise ValueError("krylov dimension too small")
    numiter = 1
    f = arnoldi_initialize(A, x0, orth)
    numops = 1
    converged = 0
    T = U = fv = None
    while True:
        beta = f.normres()
        K = f.k
        if K == krylovdim or beta <= tol or (eager and K >= howmany):
            H = arnoldi_rayleighquotient(f)
            T, U, values = hschur(H)
            p = eigsort_complex(which)(values)
            T, U, values = permuteschur(T, U, p)
            fv = U[K - 1, :] * beta
            converged = 0
            while converged < K and abs(fv[converged]) <= tol:
                converged += 1
            if 0 < converged < K and T[converged, converged - 1] != 0:
                converged -= 1
            if converged >= howmany or beta <= tol:
                break
        if K < krylovdim:
            f = arnoldi_expand(A, f, orth)
            numops += 1
        else:
            if numiter == maxiter:
                break
            keep = (3 * krylovdim + 2 * converged) // 5
            H = np.array(T)
            if H[keep, keep - 1] != 0:                 # in the middle of a 2×2 block
                if keep > 1:
                    keep -= 1
                else:
                    keep += 1
                    if krylovdim == 2:
                        break
            restore_arnoldi_form(U, H, fv, keep)
            # copy H back into packed Hessenberg storage (only the first keep columns survive the shrink)
            for j in range(1, K + 1):
                for i in range(1, min(j + 1, K) + 1):
                    f.H[hidx(i, j)] = H[i - 1, j - 1]
            basistransform(f.V, U[:, :keep])
            f.V[keep] = f.r * (1 / beta)
            f = arnoldi_shrink(f, keep)
            numiter += 1
    return T, U, f, converged, numiter, numops


def _schur_howmany(T, f, howmany, converged):
    hm = howmany
    if howmany < f.k and T[howmany, howmany - 1] != 0:
        hm += 1
    elif T.shape[0] < howmany:
        hm = T.shape[0]
    if converged > howmany:
        hm = converged
    return hm


def schursolve_arnoldi(A, x0, howmany, which, krylovdim=30, maxiter=100, tol=1e-12,
                       orth: Orth = Orth(MGS2), eager=False):
    """schursolve(A, x₀, howmany, which, ::Arnoldi) — eigsolve/arnoldi.jl:110-145."""
    T, U, f, converged, numiter, numops = _schursolve(A, x0, howmany, which, krylovdim, maxiter, tol, orth, eager)
    hm = _schur_howmany(T, f, howmany, converged)
    TT = T[:hm, :hm]
    values = schur2eigvals(TT)
    vectors = [unproject(np.zeros_like(f.V[0]), f.V, U[:, i]) for i in range(hm)]
    residuals = [f.r * U[-1, i] for i in range(hm)]
    normres = np.array([f.normres() * abs(U[-1, i]) for i in range(hm)])
    return TT, vectors, values, dict(converged=converged, residual=residuals, normres=normres,
                                     numiter=numiter, numops=numops)


def eigsolve_arnoldi(A, x0, howmany, which, krylovdim=30, maxiter=100, tol=1e-12,
                     orth: Orth = Orth(MGS2), eager=False):
    """eigsolve(A, x₀, howmany, which, ::Arnoldi) — eigsolve/arnoldi.jl:147-184.  Complex eigenvectors."""
    T, U, f, converged, numiter, numops = _schursolve(A, x0, howmany, which, krylovdim, maxiter, tol, orth, eager)
    hm = _schur_howmany(T, f, howmany, converged)
    TT = T[:hm, :hm]
    values = schur2eigvals(TT)
    Vc = U[:, :hm] @ schur2eigvecs(TT)
    Bm = np.column_stack(f.V)
    vectors = [Bm @ Vc[:, i] for i in range(hm)]
    residuals = [f.r * Vc[-1, i] for i in range(hm)]
    normres = np.array([f.normres() * abs(Vc[-1, i]) for i in range(hm)])
    return values, vectors, dict(converged=converged, residual=residuals, normres=normres,
                                 numiter=numiter, numops=numops)


def schur2realeigvecs(T):
    """schur2realeigvecs — dense/linalg.jl:247-257: eigenvectors of an upper TRIANGULAR real matrix
    (back substitution), unit 2-norm columns."""
    n = T.shape[0]
    if np.any(np.diag(T, -1) != 0):
        raise ValueError("T must be upper triangular")
    V = np.zeros((n, n))
    smin = np.finfo(float).eps * max(np.abs(T).max(), 1.0)
    for k in range(n):
        V[k, k] = 1.0
        for i in range(k - 1, -1, -1):
            den = T[i, i] - T[k, k]
            if abs(den) < smin:
                den = smin
            V[i, k] = -(T[i, i + 1:k + 1] @ V[i + 1:k + 1, k]) / den
        V[:, k] /= np.linalg.norm(V[:, k])
    return V


def realeigsolve_arnoldi(A, x0, howmany, which, krylovdim=30, maxiter=100, tol=1e-12,
                         orth: Orth = Orth(MGS2), eager=False):
    """realeigsolve(A, x₀, howmany, which, ::Arnoldi) — eigsolve/arnoldi.jl:293-349: the caller asserts a
    real spectrum; 2×2 Schur blocks among the requested values are flattened (their imaginary parts
    dropped), real eigenvalues and real eigenvectors are returned."""
    T, U, f, converged, numiter, numops = _schursolve(A, x0, howmany, which, krylovdim, maxiter, tol, orth, eager)
    T = np.array(T)
    ignored = []
    i = 0
    while i < howmany:
        i += 1
        if i < f.k:
            if abs(T[i, i - 1]) > tol:
                ignored.append(math.sqrt(max(-T[i, i - 1] * T[i - 1, i], 0.0)))
            T[i, i - 1] = 0
    while i < converged:
        i += 1
        if i < f.k:
            if abs(T[i, i - 1]) <= tol:
                T[i, i - 1] = 0
            else:
                i -= 1
                break
    hm = min(i, T.shape[0])
    converged = min(converged, hm)
    TT = T[:hm, :hm]
    values = np.diag(TT).copy()
    Vr = U[:, :hm] @ schur2realeigvecs(TT)
    vectors = [unproject(np.zeros_like(f.V[0]), f.V, Vr[:, j]) for j in range(hm)]
    residuals = [f.r * Vr[-1, j] for j in range(hm)]
    normres = np.array([f.normres() * abs(Vr[-1, j]) for j in range(hm)])
    return values, vectors, dict(converged=converged, residual=residuals, normres=normres, numiter=numiter,
                                 numops=numops, ignored_imag=ignored)


def linsolve_gmres(A, b, x0=None, krylovdim=30, maxiter=100, tol=1e-12, orth: Orth = Orth(MGS2),
                   a0=0.0, a1=1.0):
    """linsolve(operator, b, x₀, ::GMRES, a₀, a₁) — src/linsolve/gmres.jl:1-151.
    tol is the absolute tolerance (linsolve.jl:159-161 resolves max(atol, rtol*‖b‖))."""
    if x0 is None:
        x0 = np.zeros_like(b)
    y0 = apply(A, x0)
    r = b * 1.0
    if a0 != 0:
        r = r - a0 * x0
    r = r - a1 * y0
    x = x0 * 1.0
    beta = norm(r)
    if beta < tol:
        return x, dict(converged=1, residual=r, normres=beta, numiter=0, numops=1)
    y = np.zeros(krylovdim + 1)
    gs = [None] * krylovdim
    R = np.zeros((krylovdim, krylovdim))
    numiter, numops = 0, 1
    f = arnoldi_initialize(A, r, orth)
    numops += 1
    while True:
        numiter += 1
        y[0] = beta
        k = 1
        R[0, 0] = a0 + a1 * f.h(1, 1)
        c, s, rr = givens(R[0, 0], a1 * f.normres())
        gs[0] = (c, s)
        R[0, 0] = rr
        y[1] = 0.0
        y[0], y[1] = c * y[0] + s * y[1], -s * y[0] + c * y[1]
        beta = abs(y[1])
        while R[k - 1, k - 1] != 0 and beta > tol and f.k < krylovdim:
            f = arnoldi_expand(A, f, orth)
            numops += 1
            k = f.k
            for i in range(1, k):
                R[i - 1, k - 1] = a1 * f.h(i, k)
            R[k - 1, k - 1] = a0 + a1 * f.h(k, k)
            for i in range(k - 1):
                c, s = gs[i]
                R[i, k - 1], R[i + 1, k - 1] = (c * R[i, k - 1] + s * R[i + 1, k - 1],
                                                -s * R[i, k - 1] + c * R[i + 1, k - 1])
            if math.hypot(R[k - 1, k - 1], a1 * f.normres()) < tol:
                c, s, rr = givens(0.0, y[k - 1])     # givens(zero, y[k], k+1, k): rotate weight into y[k+1]
                gs[k - 1] = ("swap", c, s)
                y[k] = rr
                y[k - 1] = 0.0
                R[k - 1, k - 1] = 0.0
            else:
                c, s, rr = givens(R[k - 1, k - 1], a1 * f.normres())
                gs[k - 1] = (c, s)
                R[k - 1, k - 1] = rr
                y[k] = 0.0
                y[k - 1], y[k] = c * y[k - 1] + s * y[k], -s * y[k - 1] + c * y[k]
            beta = abs(y[k])
        if R[k - 1, k - 1] == 0 and y[k - 1] == 0:
            ldiv_upper(R, y, k - 1)
        else:
            ldiv_upper(R, y, k)
        for i in range(k):
            x = x + y[i] * f.V[i]
        if beta > tol and numiter < maxiter:
            # residual without a new operator application (gmres.jl:110-117): rotate the
            # whole basis with the adjoint Givens sequence, take column k+1
            V = list(f.V) + [f.r * (1 / f.normres())]
            for i in range(k):
                g = gs[i]
                if g[0] == "swap":                         # Givens(k+1, k, c, s) of the singular branch, gmres.jl:84
                    givens_rmul_basis(V, i + 1, i, g[1], -g[2])
                    continue
                c, s = g
                givens_rmul_basis(V, i, i + 1, c, -s)      # rmul!(V, gs[i]')
            r = V[k] * y[k]
        else:
            r = b * 1.0
            r = r - apply(A, x, a0, a1)
            numops += 1
            beta = norm(r)
            if beta < tol:
                return x, dict(converged=1, residual=r, normres=beta, numiter=numiter, numops=numops)
        if numiter >= maxiter:
            return x, dict(converged=0, residual=r, normres=beta, numiter=numiter, numops=numops)
        f = arnoldi_initialize_inplace(A, r, f, orth)


def linsolve_cg(A, b, x0=None, maxiter=100, tol=1e-12, a0=0.0, a1=1.0):
    """linsolve(operator, b, x₀, ::CG, a₀, a₁) — src/linsolve/cg.jl:1-103."""
    if x0 is None:
        x0 = np.zeros_like(b)
    y0 = apply(A, x0)
    r = b * 1.0
    if a0 != 0:
        r = r - a0 * x0
    r = r - a1 * y0
    x = x0 * 1.0
    normr = norm(r)
    numops, numiter = 1, 0
    if normr < tol:
        return x, dict(converged=1, residual=r, normres=normr, numiter=numiter, numops=numops)
    rho = normr * normr       # Julia literal_pow: normr^2 = normr*normr
    p = r * 1.0
    q = apply(A, p, a0, a1)
    alpha = rho / inner(p, q)
    x = x + alpha * p
    r = r - alpha * q
    normr = norm(r)
    rhoold = rho
    rho = normr * normr       # Julia literal_pow: normr^2 = normr*normr
    beta = rho / rhoold
    numops += 1
    numiter += 1
    if normr < tol:
        return x, dict(converged=1, residual=r, normres=normr, numiter=numiter, numops=numops)
    while True:
        p = beta * p + r
        q = apply(A, p, a0, a1)
        alpha = rho / inner(p, q)
        x = x + alpha * p
        r = r - alpha * q
        normr = norm(r)
        if normr < tol:
            r = b - apply(A, x, a0, a1)
            normr = norm(r)
            rho = normr * normr       # Julia literal_pow: normr^2 = normr*normr
            beta = 0.0
        else:
            rhoold = rho
            rho = normr * normr       # Julia literal_pow: normr^2 = normr*normr
            beta = rho / rhoold
        numops += 1
        numiter += 1
        if normr < tol:
            return x, dict(converged=1, residual=r, normres=normr, numiter=numiter, numops=numops)
        if numiter >= maxiter:
            return x, dict(converged=0, residual=r, normres=normr, numiter=numiter, numops=numops)


def linsolve_bicgstab(A, b, x0=None, maxiter=100, tol=1e-12, a0=0.0, a1=1.0):
    """linsolve(operator, b, x₀, ::BiCGStab, a₀, a₁) — src/linsolve/bicgstab.jl:1-203.
    The reference unrolls the first iteration (:32-93) ahead of its `while true` (:95-201);
    the two bodies differ only in how p is formed, so one loop with a flag restates both."""
    if x0 is None:
        x0 = np.zeros_like(b)
    info = lambda c, res, nr, it, ops: dict(converged=c, residual=res, normres=nr, numiter=it, numops=ops)
    y0 = apply(A, x0)
    r = b * 1.0                                        # :8-10
    if a0 != 0:
        r = r - a0 * x0
    r = r - a1 * y0
    x = x0 * 1.0
    normr = norm(r)
    numops, numiter = 1, 0
    if normr < tol:                                    # :22
        return x, info(1, r, normr, numiter, numops)
    r_shadow = r * 1.0                                 # :32
    rho = alpha = omega = 1.0
    p = v = None
    while True:
        numiter += 1
        rhoold, rho = rho, inner(r_shadow, r)          # :33 / :97-98
        if p is None:
            if rho == 0.0:                             # `ρ ≈ 0.0` with default tolerances is ρ == 0 (:36)
                return x, info(0, r, normr, numiter, numops)
            p = r * 1.0                                # :44
        else:
            beta = (rho / rhoold) * (alpha / omega)    # :99
            p = p - omega * v                          # :101-102
            p = beta * p + r
        v = apply(A, p, a0, a1)                        # :45 / :104
        numops += 1
        sigma = inner(r_shadow, v)
        alpha = rho / sigma
        s = r - alpha * v                              # half step residual
        xhalf = x + alpha * p                          # half step iterate
        normr = norm(s)
        if normr < tol:                                # :60-72 / :118-135 — check the ACTUAL residual
            s = b - apply(A, xhalf, a0, a1)            # s is overwritten whether or not it passes
            numops += 1
            normr_act = norm(s)
            if normr_act < tol:
                return xhalf, info(1, s, normr_act, numiter, numops)
        t = apply(A, s, a0, a1)                        # :75 / :141
        numops += 1
        omega = inner(t, s) / inner(t, t)
        x = xhalf + omega * s                          # full step
        r = s - omega * t
        normr = norm(r)
        if normr < tol:                                # :87-100 / :153-169
            r = b - apply(A, x, a0, a1)                # r is overwritten whether or not it passes
            numops += 1
            normr_act = norm(r)
            if normr_act < tol:
                return x, info(1, r, normr_act, numiter, numops)
        if numiter > 1 and numiter >= maxiter:         # :170 — inside the while loop only
            return x, info(0, r, normr, numiter, numops)


def lssolve_lsmr(A, b, maxiter=100, krylovdim=30, tol=1e-12, orth: "Orth | None" = None, lam=0.0):
    """lssolve(operator, b, ::LSMR, λ) — src/lssolve/lsmr.jl:1-162.  Minimises
    ‖b − A x‖² + λ²‖x‖² by Golub-Kahan bidiagonalisation with two layers of plane rotations; the last
    `krylovdim` right vectors are kept in a ring (slot order, not age, defines the sweep order) and v
    is reorthogonalised against them (:75-88).  info.normres = |ζ̄| estimates ‖Aᴴ r − λ² x‖."""
    if orth is None:
        orth = Orth(MGS)                               # algorithms.jl:517
    info = lambda c, res, nr, it, ops: dict(converged=c, residual=res, normres=nr, numiter=it, numops=ops)
    u = b * 1.0
    v = apply_adjoint(A, b) * 1.0
    beta = norm(u)
    u = u / beta
    v = v / beta
    alpha = norm(v)
    v = v / alpha
    V = [v]
    K = krylovdim
    Vv = np.zeros(K)
    alphabar, zetabar = alpha, alpha * beta
    rho, theta, rhobar, cbar, sbar = 1.0, 0.0, 1.0, 1.0, 0.0
    abszetabar = abs(zetabar)
    x = np.zeros_like(v)
    h = v * 1.0
    hbar = np.zeros_like(v)
    r = u * beta
    Ah = np.zeros_like(u)
    Ahbar = np.zeros_like(u)
    numiter, numops = 0, 1
    if abszetabar < tol:
        return x, info(1, r, abszetabar, numiter, numops)
    while True:
        numiter += 1
        Av = apply_normal(A, v)
        numops += 1
        Ah = Av - (theta / rho) * Ah                   # :64  Ah ← Av − θ/ρ·Ah
        u = Av - alpha * u                             # :67  (Av is consumed)
        beta = norm(u)
        if beta > tol:
            u = u / beta
            v = apply_adjoint(A, u) - beta * v         # :72
            numops += 1
            if K > 1:
                v, _ = orthogonalize(v, V, Vv[:min(K, numiter)], orth)
            alpha = norm(v)
            if alpha > tol:
                v = v / alpha
                if numiter < K:
                    V.append(v)
                else:
                    V[numiter % K] = v                 # mod1(numiter + 1, K), 1-based
        alphahat = math.hypot(alphabar, lam)           # rotation P̂ (regularisation)
        rhoold = rho
        rho = math.hypot(alphahat, beta)               # rotation P: B → R
        c, s_ = alphahat / rho, beta / rho
        theta = s_ * alpha
        alphabar = c * alpha
        rhobarold = rhobar                             # rotation P̄: Rᵀ → R̄
        thetabar = sbar * rho
        cbarrho = cbar * rho
        rhobar = math.hypot(cbarrho, theta)
        cbar = cbarrho / rhobar
        sbar = theta / rhobar
        zeta = cbar * zetabar
        zetabar = -sbar * zetabar
        g = -thetabar * rho / (rhoold * rhobarold)
        hbar = h + g * hbar
        Ahbar = Ah + g * Ahbar
        x = x + (zeta / (rho * rhobar)) * hbar
        r = r - (zeta / (rho * rhobar)) * Ahbar
        h = v - (theta / rho) * h
        abszetabar = abs(zetabar)
        if abszetabar <= tol:
            return x, info(1, r, abszetabar, numiter, numops)
        if numiter >= maxiter:
            return x, info(0, r, abszetabar, numiter, numops)


# ------------------------------------------------------------------ expintegrator ---------

def lanczos_initialize_inplace(A, x0, f: LanczosFact, orth: Orth):
    """initialize!(iter, state) — factorizations/lanczos.jl:223-249."""
    while len(f.V) > 1:
        f.V.pop()
    f.V[0] = x0 * (1 / norm(x0))
    w = apply(A, f.V[0])
    r, alpha = orthogonalize_vec(w, f.V[0], orth)
    f.k = 1
    f.alphas[:] = [alpha]
    f.betas[:] = [norm(r)]
    f.r = r
    return f


def expintegrator(A, t, u, method="lanczos", orth: Orth = Orth(MGS2), krylovdim=30, maxiter=100, tol=1e-12,
                  eager=False):
    """expintegrator(A, t, u::Tuple, alg) — matrixfun/expintegrator.jl:101-323, real t.
    y(t) = ϕ₀(tA) u₀ + t ϕ₁(tA) u₁ + … + tᵖ ϕₚ(tA) uₚ, the solution of y' = A y + Σ_j t^j/j! u_{j+1},
    by adaptive Krylov time stepping (Niesen & Wright).  `tol` is the error per unit time."""
    from scipy.linalg import expm
    u = tuple(u)
    if len(u) == 1:
        u = (u[0], np.zeros_like(u[0]))
    p = len(u) - 1
    lanczos = method == "lanczos"
    u0 = u[0]
    Au0 = apply(A, u0)
    numops = 1
    w0 = u0 * 1.0
    K = krylovdim
    eta = tol
    totalerr = 0.0
    sgn = float(np.sign(t))
    tau = abs(t)
    if math.isfinite(tau):
        dtau, dtaumin, maxerr = tau, tau / maxiter, tau * eta
    else:
        dtau, dtaumin, maxerr = 1.0, 0.0, eta
    delta, gamma = 1.2, 0.8
    tau0 = 0.0
    w = [None] * (p + 1)
    w[0] = w0
    w[1] = Au0 * 1.0

    def fill_w(first_apply_done):
        nonlocal numops
        for j in range(1, p + 1):
            if j > 1 or not first_apply_done:
                w[j] = apply(A, w[j - 1])
                numops += 1
            lfac = 1
            for l in range(0, p - j + 1):
                w[j] = w[j] + ((sgn * tau0) ** l / lfac) * u[j + l]
                lfac *= l + 1

    fill_w(True)
    beta = norm(w[p])
    if beta < eta and p == 1:
        return w0, dict(converged=1, residual=None, normres=beta, numiter=0, numops=numops)
    if lanczos:
        f = lanczos_initialize(A, w[p], orth)
        rq = lambda: (np.diag(f.alphas) + np.diag(f.betas[:f.k - 1], 1) + np.diag(f.betas[:f.k - 1], -1))
        init_inplace = lambda x: lanczos_initialize_inplace(A, x, f, orth)
        expand = lambda: lanczos_expand(A, f, orth)
    else:
        f = arnoldi_initialize(A, w[p], orth)
        rq = lambda: arnoldi_rayleighquotient(f)
        init_inplace = lambda x: arnoldi_initialize_inplace(A, x, f, orth)
        expand = lambda: arnoldi_expand(A, f, orth)
    numops += 1
    numiter = 1

    def small_exp(K, dt):
        H = np.zeros((K + p + 1, K + p + 1))
        H[:K, :K] = rq() * (sgn * dt)
        H[0, K] = 1
        for i in range(1, p + 1):
            H[K + i - 1, K + i] = 1
        return expm(H)

    def take_step(K, dt, expH):
        nonlocal w0
        jfac = 1
        for j in range(1, p):
            w0 = w0 + ((sgn * dt) ** j / jfac) * w[j]
            jfac *= j + 1
        wp = unproject(np.zeros_like(w[p]), f.V, expH[:K, K + p - 1])
        wp = wp + expH[K - 1, K + p] * f.r
        w[p] = wp
        w0 = w0 + (beta * (sgn * dt) ** p) * wp
        w[0] = w0

    while True:
        K = f.k
        if K == krylovdim:
            if numiter < maxiter:
                dtau = min(dtau, tau - tau0)
                if math.isfinite(tau):
                    dtaumin = (tau - tau0) / (maxiter - numiter + 1)
            else:
                dtau = tau - tau0
            expH = small_exp(K, dtau)
            eps_ = abs(dtau ** p * beta * f.normres() * expH[K - 1, K + p])
            omega = eps_ / (dtau * eta)
            q = K / 2
            while numiter < maxiter and omega >= 1 and dtau > dtaumin:
                eps_prev, dtau_prev = eps_, dtau
                dtau = max(dtau * (gamma / omega) ** (1 / (q + 1)), dtaumin)
                expH = small_exp(K, dtau)
                eps_ = abs(dtau ** p * beta * f.normres() * expH[K - 1, K + p])
                omega = eps_ / (dtau * eta)
                q = max(0.0, math.log(eps_ / eps_prev) / math.log(dtau / dtau_prev) - 1)
            tau0 = tau0 + dtau if numiter < maxiter else tau
            totalerr += eps_
            take_step(K, dtau, expH)
            if omega < gamma:
                dtau *= (gamma / omega) ** (1 / (q + 1))
        elif f.normres() <= (tau - tau0) * eta or eager:
            dt = tau - tau0
            expH = small_exp(K, dt)
            eps_ = abs(dt ** p * beta * f.normres() * expH[K - 1, K + p])
            omega = eps_ / (dt * eta)
            if omega < 1:
                totalerr += eps_
                take_step(K, dt, expH)
                tau0 = tau
        if tau0 >= tau:
            return w0, dict(converged=int(totalerr <= maxerr), residual=None, normres=totalerr,
                            numiter=numiter, numops=numops)
        if K < krylovdim:
            expand()
            numops += 1
        else:
            fill_w(False)
            beta = norm(w[p])
            if beta < eta and p == 1:
                return w0, dict(converged=1, residual=None, normres=beta, numiter=numiter, numops=numops)
            init_inplace(w[p])
            numops += 1
            numiter += 1


# ------------------------------------------------------------------ GKL / svdsolve --------

@dataclass
class GKLFact:
    k: int
    U: list
    V: list
    alphas: list
    betas: list
    r: np.ndarray

    def normres(self):
        return self.betas[-1]


def gkl_initialize(A, u0, orth: Orth):
    """initialize(iter::GKLIterator) — gkl.jl:183-215."""
    beta0 = norm(u0)
    if beta0 == 0:
        raise ValueError("initial vector should not have norm zero")
    v0 = apply_adjoint(A, u0)
    alpha = norm(v0) / beta0
    Av0 = apply_normal(A, v0)
    alpha2 = inner(u0, Av0) / (beta0 * beta0)
    if not np.isclose(alpha2, alpha * alpha, rtol=math.sqrt(np.finfo(np.asarray(u0).dtype).eps)):
        raise ValueError("operator and its adjoint are not compatible")
    T = np.asarray(u0).dtype.type
    u = u0 * T(1 / beta0)
    v = v0 * T(1 / (alpha * beta0))
    r = Av0 * T(1 / (alpha * beta0))
    r = r - T(alpha) * u
    beta = norm(r)
    return GKLFact(1, [u], [v], [alpha], [beta], r)


def gkl_recurrence(A, U, V, beta, orth: Orth):
    """gklrecurrence — gkl.jl:294-404."""
    t = orth.tag
    T = U[-1].dtype.type
    eps = np.finfo(U[-1].dtype).eps
    u = U[-1]
    v = apply_adjoint(A, u)
    v = v - T(beta) * V[-1]
    if t == MGS2:
        for q in V:
            s = inner(q, v)
            v = v - T(s) * q
    alpha = norm(v)
    if t in (CGSIR, MGSIR):
        nold = math.sqrt(alpha * alpha + beta * beta)
        while (alpha < orth.eta * nold) if t == CGSIR else (eps < alpha < orth.eta * nold):
            nold = alpha
            if t == CGSIR:
                s = np.empty(len(V))
                v, s = _cgs_pass(v, V, s)
            else:
                for q in V:
                    s = inner(q, v)
                    v = v - T(s) * q
            alpha = norm(v)
    v = v * T(1 / alpha)
    r = apply_normal(A, v)
    r = r - T(alpha) * u
    if t == CGS2:
        s = np.empty(len(U))
        r, s = _cgs_pass(r, U, s)
    elif t == MGS2:
        for q in U:
            s = inner(q, r)
            r = r - T(s) * q
    beta = norm(r)
    if t in (CGSIR, MGSIR):
        nold = math.sqrt(alpha * alpha + beta * beta)
        while eps < beta < orth.eta * nold:
            nold = beta
            if t == CGSIR:
                s = np.empty(len(U))
                r, s = _cgs_pass(r, U, s)
            else:
                for q in U:
                    s = inner(q, r)
                    r = r - T(s) * q
            beta = norm(r)
    return v, r, alpha, beta


def gkl_expand(A, f: GKLFact, orth: Orth):
    """expand! — gkl.jl:246-269."""
    betaold = f.normres()
    T = f.r.dtype.type
    f.U.append(f.r * T(1 / betaold))
    v, r, alpha, beta = gkl_recurrence(A, f.U, f.V, betaold, orth)
    f.V.append(v)
    f.alphas.append(alpha)
    f.betas.append(beta)
    f.k += 1
    f.r = r
    return f


def gkl_shrink(f: GKLFact, k):
    """shrink! — gkl.jl:270-291."""
    if f.k <= k:
        return f
    while len(f.V) > k + 1:
        f.U.pop()
        f.V.pop()
    f.V.pop()
    r = f.U.pop()
    del f.alphas[k:]
    del f.betas[k:]
    f.k = k
    f.r = r * r.dtype.type(f.normres())
    return f


def bidiagsvd_lower(alphas, betas):
    """bidiagsvd!(Bidiagonal(αs, βs, :L)) -> (P, S, Q) with B = P*Diag(S)*Q (Q = Vt).
    The reference calls LAPACK bdsqr (dense/linalg.jl:123-130); values identical,
    vectors up to a common sign per triplet."""
    K = len(alphas)
    B = np.diag(np.asarray(alphas, dtype=np.float64))
    for i in range(K - 1):
        B[i + 1, i] = betas[i]
    P, S, Q = np.linalg.svd(B)
    return P, S, Q


def svdsolve_gkl(A, u0, howmany=1, which="LR", krylovdim=30, maxiter=100, tol=1e-12,
                 orth: Orth = Orth(MGS2), eager=False):
    """svdsolve(A, x₀, howmany, which, ::GKL) — src/eigsolve/svdsolve.jl:144-314."""
    if howmany > krylovdim:
        raise ValueError("krylov dimension too small")
    numiter = 1
    f = gkl_initialize(A, u0, orth)
    numops = 2
    HH = np.zeros((krylovdim + 1, krylovdim))
    converged = 0
    P = S = Q = fvec = None
    while True:
        beta = f.normres()
        K = f.k
        if K == krylovdim or beta <= tol or (eager and K >= howmany):
            P, S, Q = bidiagsvd_lower(f.alphas[:K], f.betas[:K - 1])
            if which == "SR":
                P = P[:, ::-1].copy()
                S = S[::-1].copy()
                Q = Q[::-1, :].copy()
            elif which != "LR":
                raise ValueError(which)
            fvec = Q.T[K - 1, :] * beta          # mul!(f, view(Q', K, :), β)
            converged = 0
            while converged < K and abs(fvec[converged]) < tol:
                converged += 1
            if converged >= howmany or beta <= tol:
                break
        if K < krylovdim:
            f = gkl_expand(A, f, orth)
            numops += 2
        else:
            if numiter == maxiter:
                break
            keep = (3 * krylovdim + 2 * converged) // 5
            T = f.r.dtype.type
            Pc = P.astype(f.r.dtype)
            Qc = Q.T.astype(f.r.dtype)
            basistransform(f.U, Pc[:, :keep])
            basistransform(f.V, Qc[:, :keep])
            f.U[keep] = f.r * T(1 / f.normres())
            H = HH[: keep + 1, :keep]
            H[:] = 0
            for j in range(keep):
                H[j, j] = S[j]
                H[keep, j] = fvec[j]
            for j in range(keep - 1, -1, -1):
                rr = list(range(0, j + 1))
                hb, hv, nu = householder_vec(H[j + 1, rr], j)
                H[j + 1, j] = nu
                H[j + 1, :j] = 0
                householder_rmul(H[: j + 1, :], hb, hv, rr)
                householder_rmul_basis(f.V, T(hb), hv.astype(f.r.dtype), rr)
                hb, hv, nu = householder_vec(H[rr, j], j)
                H[j, j] = nu
                H[:j, j] = 0
                householder_lmul(hb, hv, rr, H[:, :j])
                householder_rmul_basis(f.U, T(hb), hv.astype(f.r.dtype), rr)
            for j in range(keep):
                f.alphas[j] = H[j, j]
                f.betas[j] = H[j + 1, j]
            f = gkl_shrink(f, keep)
            numiter += 1
    if converged > howmany:
        howmany = converged
    values = S[:howmany].copy()
    dt = f.r.dtype
    left = [unproject(np.zeros_like(f.U[0]), f.U, P[:, i].astype(dt)) for i in range(howmany)]
    right = [unproject(np.zeros_like(f.V[0]), f.V, Q[i, :].astype(dt)) for i in range(howmany)]
    residuals = [f.r * dt.type(Q[i, -1]) for i in range(howmany)]
    normres = np.abs(fvec[:howmany])
    return values, left, right, dict(converged=converged, residual=residuals, normres=normres,
                                     numiter=numiter, numops=numops, fact=f)


# ------------------------------------------------------------------ block primitives ------

def block_inner(X, Y):
    """blocklanczos.jl:43-52."""
    M = np.empty((len(X), len(Y)))
    for j in range(len(Y)):
        for i in range(len(X)):
            M[i, j] = inner(X[i], Y[j])
    return M


def block_reorthogonalize(R, V):
    """blocklanczos.jl:277-284."""
    for i in range(len(R)):
        for q in V:
            s = inner(q, R[i])
            R[i] = R[i] - s * q
    return R


def block_qr(block, tol):
    """block_qr! — blocklanczos.jl:312-353.  Returns (R[good,:], good_idx, is_drift)."""
    n = len(block)
    drift = False
    idx = [True] * n
    R = np.zeros((n, n))
    beta = math.sqrt(inner(block[0], block[0]))
    if beta > tol:
        R[0, 0] = beta
        block[0] = block[0] * (1 / beta)
    else:
        block[0] = block[0] * 0.0
        idx[0] = False
    for j in range(1, n):
        for i in range(j):
            R[i, j] = inner(block[i], block[j])
            block[j] = block[j] - R[i, j] * block[i]
        beta = norm(block[j])
        if tol < beta < 100 * tol:
            drift = True
            for i in range(j):
                d = inner(block[i], block[j])
                R[i, j] += d
                block[j] = block[j] - d * block[i]
            beta = norm(block[j])
        if beta < tol:
            block[j] = block[j] * 0.0
            idx[j] = False
        else:
            R[j, j] = beta
            block[j] = block[j] * (1 / beta)
    good = [i for i in range(n) if idx[i]]
    return R[good, :], good, drift


# ------------------------------------------------------------------ block Lanczos ---------

@dataclass
class BlockLanczosFact:
    """BlockLanczosFactorization — blocklanczos.jl:82-94.  A V = V H + R Bᵀ, B = [0; I]."""
    k: int
    V: list
    H: np.ndarray
    R: list
    R_size: int
    norm_R: float


def _block_norm(R):
    return math.sqrt(sum(inner(x, x) for x in R))


def blocklanczos_initialize(A, X0, maxdim, qr_tol):
    """initialize(::BlockLanczosIterator) — blocklanczos.jl:159-190."""
    if _block_norm(X0) == 0:
        raise ValueError("initial vector should not have norm zero")
    X1 = [x * 1.0 for x in X0]
    _, good, _ = block_qr(X1, qr_tol)
    X1 = [X1[i] for i in good]
    V = list(X1)
    bs = len(X1)
    AX = [apply(A, x) for x in X1]
    M = block_inner(X1, AX)
    H = np.zeros((maxdim, maxdim))
    H[:bs, :bs] = M
    for j in range(bs):
        for i in range(bs):
            AX[j] = AX[j] - M[i, j] * X1[i]
    return BlockLanczosFact(bs, V, H, AX, bs, _block_norm(AX))


def blocklanczos_recurrence(A, V, B):
    """block_lanczosrecurrence(…, ::ModifiedGramSchmidt2) — blocklanczos.jl:232-251."""
    bs, bs_prev = B.shape
    k = len(V)
    X = V[k - bs:k]
    AX = [apply(A, x) for x in X]
    M = block_inner(X, AX)
    Xprev = V[k - bs_prev - bs:k - bs]
    for j in range(bs):
        for i in range(bs):
            AX[j] = AX[j] - M[i, j] * X[i]
        for i in range(len(Xprev)):
            AX[j] = AX[j] - B[j, i] * Xprev[i]
    block_reorthogonalize(AX, V)
    return AX, M


def blocklanczos_expand(A, f: BlockLanczosFact, qr_tol):
    """expand!(::BlockLanczosIterator, state) — blocklanczos.jl:192-230."""
    k = f.k
    R = f.R[:f.R_size]
    bs = len(R)
    Rcopy = [x * 1.0 for x in R]
    B, good, drift = block_qr(R, qr_tol)
    if drift:
        block_reorthogonalize(R, f.V)
        _, good, drift = block_qr(R, qr_tol)
        B = block_inner([R[i] for i in good], Rcopy)
    bs_next = len(good)
    f.V.extend(R[i] for i in good)
    f.H[k:k + bs_next, k - bs:k] = B[:bs_next, :bs]
    f.H[k - bs:k, k:k + bs_next] = B[:bs_next, :bs].T
    Rnext, M = blocklanczos_recurrence(A, f.V, B)
    f.H[k:k + bs_next, k:k + bs_next] = M[:bs_next, :bs_next]
    f.R[:bs_next] = Rnext
    f.norm_R = _block_norm(Rnext)
    f.k += bs_next
    f.R_size = bs_next
    return f


def eigsolve_blocklanczos(A, X0, howmany, which, krylovdim=100, maxiter=100, tol=1e-12, qr_tol=1e-12,
                          eager=False):
    """eigsolve(A, x₀::Block, howmany, which, ::BlockLanczos) — eigsolve/blocklanczos.jl:1-144."""
    if howmany > krylovdim:
        raise ValueError("krylov dimension too small")
    bs = len(X0)
    f = blocklanczos_initialize(A, X0, krylovdim + bs, qr_tol)
    numops, numiter, converged = bs + 1, 1, 0
    D = U = normres = None
    while True:
        K, beta = f.k, f.norm_R
        if K >= krylovdim or beta <= tol or (eager and K >= howmany):
            BTD = f.H[:K, :K]
            D, U = np.linalg.eigh((BTD + BTD.T) / 2)
            D, U = permuteeig(D, U, eigsort(which)(D))
            bs_R = f.R_size
            r = f.R[:bs_R]
            UU = U[K - bs_R:K, :]
            RR = block_inner(r, r)
            normres = np.sqrt(np.maximum(np.einsum("ik,ij,jk->k", UU, RR, UU), 0.0))
            converged = 0
            while converged < K and normres[converged] <= tol:
                converged += 1
            if converged >= howmany or beta <= tol:
                break
        if K < krylovdim:
            blocklanczos_expand(A, f, qr_tol)
            numops += f.R_size
        else:
            if numiter >= maxiter:
                break
            keep = max((3 * krylovdim + 2 * converged) // (5 * bs), 1) * bs
            H = np.zeros((keep + bs, keep))
            for j in range(keep):
                H[j, j] = D[j]
                H[keep:, j] = U[K - bs:K, j]
            for j in range(keep - 1, -1, -1):                 # Julia j = keep:-1:1
                rr = list(range(j + 1))
                row = j + bs                                  # Julia row j + bs
                hb, hv, nu = householder_vec(H[row, rr], j)
                H[row, j] = nu
                H[row, :j] = 0
                householder_lmul(hb, hv, rr, H)
                householder_rmul(H[:row, :], hb, hv, rr)
                householder_rmul(U, hb, hv, rr)
            f.H[:] = 0
            Hk = H[:keep, :keep]
            f.H[:keep, :keep] = (Hk + Hk.T) / 2
            basistransform(f.V, U[:, :keep])
            view_H = H[keep + bs - bs_R:keep + bs, keep - bs_R:keep]
            Rnew = list(f.R[:bs_R])
            basistransform(Rnew, view_H)
            f.R[:bs_R] = Rnew[:bs_R]
            del f.V[keep:]
            f.k = keep
            numiter += 1
    hm = howmany
    if converged > howmany:
        hm = converged
    elif len(D) < howmany:
        hm = len(D)
    values = D[:hm].copy()
    vectors = [unproject(np.zeros_like(f.V[0]), f.V, U[:, i]) for i in range(hm)]
    bs_R, K = f.R_size, f.k
    U2 = U[K - bs_R:K, :hm]
    residuals = []
    for i in range(hm):
        res = np.zeros_like(f.R[0])
        for j in range(bs_R):
            res = res + U2[j, i] * f.R[j]
        residuals.append(res)
    return values, vectors, dict(converged=converged, residual=residuals, normres=normres[:hm],
                                 numiter=numiter, numops=numops, fact=f)


# ------------------------------------------------------------------ fixtures --------------

def toric_code_hamiltonian(m, n):
    """test/eigsolve.jl:471-533: H = Σ X-plaquettes + Σ Z-vertices (last of each dropped)."""
    N = 2 * m * n

    def li(i, j):       # LinearIndices((m, n))[mod1(i,m), mod1(j,n)], 1-based
        return ((j - 1) % n) * m + ((i - 1) % m) + 1

    def bottom(i, j):
        return li(i, j) + m * n

    def right(i, j):
        return li(i, j)

    xs, zs = [], []
    for i in range(1, m + 1):        # Julia: for i in 1:m, j in 1:n  (j inner)
        for j in range(1, n + 1):
            xs.append((bottom(i, j + 1), right(i, j), bottom(i, j), right(i - 1, j)))
            zs.append((right(i, j), bottom(i, j), right(i, j - 1), bottom(i + 1, j)))
    dim = 2 ** N
    idx = np.arange(dim, dtype=np.int64)
    rows, cols, vals = [], [], []
    diag = np.zeros(dim)
    for s in xs[:-1]:
        mask = 0
        for pos in s:                # qubit `pos` (1-based, kron order: pos 1 = most significant)
            mask ^= 1 << (N - pos)
        rows.append(idx)
        cols.append(idx ^ mask)
        vals.append(np.ones(dim))
    for s in zs[:-1]:
        sign = np.ones(dim)
        for pos in s:
            sign *= 1 - 2 * ((idx >> (N - pos)) & 1)
        diag += sign
    rows.append(idx)
    cols.append(idx)
    vals.append(diag)
    H = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(dim, dim))
    H.sum_duplicates()
    H.sort_indices()
    return H
