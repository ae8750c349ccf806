"""eigsolve with the Lanczos (Krylov-Schur / thick restart) algorithm — mirror of
src/eigsolve/lanczos.jl and the user-facing entry points of src/eigsolve/eigsolve.jl.

The host keeps the restart / deflation / convergence logic exactly as KrylovKit does; all
n-length work is delegated to the device through the factorization and basis mirrors.
"""
from __future__ import annotations

import ctypes as C
import warnings

import numpy as np

from . import _lib as L

from .algorithms import Arnoldi, BlockLanczos, ConvergenceInfo, Lanczos, WARN_LEVEL
from .dense import (eigsort, householder_row, lmul_householder, permuteeig, rmul_householder,
                    tridiageigh)
from .factorizations import blocklanczos as blz
from .factorizations import lanczos as lz
from .operators import B200CSR
from .orthonormal import OrthonormalBasis, basistransform_
from .vectors import B200Context, B200Vec


def eigsolve(A, x0=None, howmany: int = 1, which: str = "LM", alg: Lanczos | None = None,
             out_vectors=None, shard=None, nccl_uid: bytes | None = None, device: int = 0, **kwargs):
    """eigsolve(A, x₀, howmany, which, alg::Lanczos) — src/eigsolve/lanczos.jl:1-155.

    A  : B200Operator / callable on B200Vec (device-resident path), or a scipy sparse
         matrix / CSR triple on the HOST (then x0 is a host array, the problem is uploaded,
         solved on the GPU and the vectors are downloaded: the end-to-end path).
    Returns (values, vectors, ConvergenceInfo).
    """
    if x0 is None:
        # eigsolve(A::AbstractMatrix, howmany, which; kwargs...) — eigsolve.jl:195-201: random start vector
        if not hasattr(A, "shape"):
            raise TypeError("eigsolve: a start vector is required unless A is a host matrix")
        x0 = np.random.default_rng().random(A.shape[0])
    if alg is None:
        alg = eigselector(A, block=isinstance(x0, blz.Block), **kwargs)
    elif kwargs:
        raise TypeError(f"eigsolve: keyword arguments {sorted(kwargs)} only apply when no algorithm is passed")
    checkwhich(which, alg)
    if isinstance(alg, BlockLanczos):
        if isinstance(x0, (list, tuple)):
            x0 = blz.Block(x0)
        return _eigsolve_blocklanczos(A, x0, howmany, which, alg)
    if isinstance(alg, Arnoldi):
        from .schursolve import eigsolve_arnoldi
        if not isinstance(x0, B200Vec):
            return _eigsolve_arnoldi_host(A, x0, howmany, which, alg, device)
        return eigsolve_arnoldi(A, x0, howmany, which, alg)
    if not isinstance(x0, B200Vec):
        return _eigsolve_host(A, x0, howmany, which, alg, out_vectors, shard, nccl_uid, device)
    return _eigsolve_lanczos(A, x0, howmany, which, alg)


def _host_issymmetric(A) -> bool:
    """LinearAlgebra.issymmetric for the host matrix types the front end accepts."""
    import scipy.sparse as sp
    if sp.issparse(A):
        return A.shape[0] == A.shape[1] and (abs(A - A.T)).nnz == 0
    if isinstance(A, np.ndarray) and A.ndim == 2:
        return A.shape[0] == A.shape[1] and bool(np.array_equal(A, A.T))
    return False


def eigselector(A, block: bool = False, issymmetric: bool | None = None, ishermitian: bool | None = None,
                krylovdim: int | None = None, maxiter: int | None = None, tol: float | None = None,
                qr_tol: float | None = None, orth=None, eager: bool = False, verbosity: int | None = None):
    """eigselector — src/eigsolve/eigsolve.jl:238-321: Lanczos for symmetric problems, Arnoldi otherwise,
    BlockLanczos for a Block start.  Symmetry is detected for host matrices (the AbstractMatrix method);
    device operators and callables default to `issymmetric = false` like a Julia function does."""
    from .algorithms import KrylovDefaults
    if issymmetric is None:
        issymmetric = _host_issymmetric(A)
    if ishermitian is None:
        ishermitian = issymmetric                      # real scalars only
    kw = dict(maxiter=KrylovDefaults.maxiter if maxiter is None else maxiter,
              tol=KrylovDefaults.tol if tol is None else tol,
              orth=KrylovDefaults.orth if orth is None else orth, eager=eager,
              verbosity=KrylovDefaults.verbosity if verbosity is None else verbosity)
    if block:
        if not (issymmetric or ishermitian):
            raise ValueError("BlockLanczos requires a symmetric or hermitian linear map. A BlockArnoldi method "
                             "has not yet been implemented")
        return BlockLanczos(krylovdim=100 if krylovdim is None else krylovdim,
                            qr_tol=KrylovDefaults.tol if qr_tol is None else qr_tol, **kw)
    kd = KrylovDefaults.krylovdim if krylovdim is None else krylovdim
    if issymmetric or ishermitian:
        return Lanczos(krylovdim=kd, **kw)
    return Arnoldi(krylovdim=kd, **kw)


def checkwhich(which: str, alg) -> None:
    """eigsolve.jl:210-222, 323-324: selector validity for the chosen algorithm (real arithmetic)."""
    from .dense import EigSorter
    if isinstance(which, EigSorter):
        if not isinstance(alg, (Lanczos, BlockLanczos)):
            probe = which.by(np.array([1j, -1j]))
            if probe[0] != probe[1]:                  # eigsolve.jl:216-221
                raise ValueError("Eigenvalue selector invalid because it does not treat `λ` and `conj(λ)` equally")
        return
    if which not in ("LM", "LR", "SR", "LI", "SI"):
        raise ValueError(f"Unknown eigenvalue selector: which = {which}")
    if which in ("LI", "SI"):
        if isinstance(alg, (Lanczos, BlockLanczos)):
            raise ValueError(f"Eigenvalue selector which = {which} invalid: real eigenvalues expected with "
                             "Lanczos and BlockLanczos algorithms")
        raise ValueError(f"Eigenvalue selector which = {which} invalid because it does not treat `λ` and "
                         "`conj(λ)` equally: that needs complex arithmetic, which this real-only engine lacks")


def _eigsolve_arnoldi_host(A, x0, howmany, which, alg, device=0):
    """Host-buffer entry for the Arnoldi driver: upload, solve, download (complex numpy vectors)."""
    import scipy.sparse as sp
    from .schursolve import eigsolve_arnoldi
    x0 = np.asarray(x0)
    if not sp.issparse(A):
        raise TypeError("eigsolve: host-side A must be a scipy sparse matrix")
    n = x0.shape[0]
    dtype = np.float32 if x0.dtype == np.float32 else np.float64
    ctx = B200Context(n, alg.krylovdim + 12, dtype=dtype, device=device)
    try:
        op = B200CSR.from_scipy(ctx, A)
        return eigsolve_arnoldi(op, ctx.from_host(x0), howmany, which, alg, to_host=True)
    finally:
        ctx.close()


USE_NATIVE_RESTART = True      # b2k_host_lanczos_restart (C++) instead of the numpy loop below
HOSTPROF: dict | None = None   # set to {} to accumulate wall seconds per driver section (bench.py reports them)


def _tick(name, t0):
    if HOSTPROF is not None:
        import time
        HOSTPROF[name] = HOSTPROF.get(name, 0.0) + (time.perf_counter() - t0)


def restart_lanczos_form(HH, D, f, U, keep, alphas, betas):
    """eigsolve/lanczos.jl:88-105 in numpy (reference implementation of the native helper)."""
    H = HH[: keep + 1, :keep]
    H[:] = 0
    for j in range(keep):
        H[j, j] = D[j]
        H[keep, j] = f[j]
    for j in range(keep - 1, -1, -1):
        h, nu = householder_row(H, j + 1, range(0, j + 1), j)
        H[j + 1, j] = nu
        H[j + 1, :j] = 0
        lmul_householder(h, H)
        rmul_householder(H, h, slice(0, j + 1))
        rmul_householder(U, h)
    for j in range(keep):
        alphas[j] = H[j, j]
        betas[j] = H[j + 1, j]


def _eigsolve_host(A, x0, howmany, which, alg, out_vectors=None, shard=None, nccl_uid=None, device=0):
    """Host-buffer entry: upload (A, x₀), solve, download.  The slab is sized for the
    factorization: krylovdim + 1 basis vectors, the residual and work columns.
    Row-sharded form (one process per GPU): `shard` = sharding.RowShard of this rank, A = the
    local rows as a CSR triple with GLOBAL column indices, x0 = the local slice."""
    import scipy.sparse as sp
    x0 = np.asarray(x0)
    n = x0.shape[0]
    dtype = np.float32 if x0.dtype == np.float32 else np.float64
    if shard is not None and shard.world > 1:
        ctx = B200Context(n, alg.krylovdim + 2 * howmany + 8, dtype=dtype, device=device, rank=shard.rank,
                          nranks=shard.world, nccl_uid=nccl_uid, n_global=shard.n_global,
                          row_offset=shard.row_offset)
        ncols_global = shard.n_global
    else:
        ctx = B200Context(n, alg.krylovdim + 2 * howmany + 8, dtype=dtype, device=device)
        ncols_global = n
    try:
        if sp.issparse(A):
            op = B200CSR.from_scipy(ctx, A)
        elif isinstance(A, tuple) and len(A) == 3:
            rowptr, colidx, vals = A
            op = B200CSR.from_csr_arrays(ctx, n, ncols_global, rowptr, colidx, vals)   # no host-side copies
        else:
            raise TypeError("eigsolve: host-side A must be a scipy sparse matrix or a CSR triple")
        xv = ctx.from_host(x0)
        vecs_h = []

        def sink(i, v):
            # out_vectors: optional preallocated (e.g. pinned) host arrays for the Ritz vectors;
            # each Ritz vector is downloaded and its slab column released before the next one
            vecs_h.append(v.to_host(out_vectors[i] if out_vectors is not None and i < len(out_vectors)
                                    else None))

        vals, _, info = _eigsolve_lanczos(op, xv, howmany, which, alg, sink)
        return vals, vecs_h, info
    finally:
        ctx.close()


def _now():
    import time
    return time.perf_counter()


FUSED_RITZ_VECTORS = True      # the hm Ritz vectors V U[:, 1:hm] in one multi-right-hand-side sweep over the basis


def _ritz_vectors(B, U: np.ndarray, hm: int):
    """[B * U[:, i] for i in 1:hm] (eigsolve/lanczos.jl:118): with the block kernels the basis is read once for all
    hm vectors instead of hm times (b2k_block_axpy on zero vectors: the same chain of fma over the basis columns)."""
    if FUSED_RITZ_VECTORS and 1 < hm <= 8 and len(B) >= 1:
        from .factorizations.blocklanczos import Block, block_axpy_
        Y = [B[0].zerovector() for _ in range(hm)]
        block_axpy_(Block(Y), B.basis, -np.asfortranarray(U[:len(B), :hm]))
        return Y
    return [B * U[:, i] for i in range(hm)]


def _eigsolve_lanczos(A, x0: B200Vec, howmany: int, which: str, alg: Lanczos, sink=None):
    krylovdim, maxiter = alg.krylovdim, alg.maxiter
    if howmany > krylovdim:
        raise ValueError(f"krylov dimension {krylovdim} too small to compute {howmany} eigenvalues")
    it = lz.LanczosIterator(A, x0, alg.orth)
    fact = lz.initialize(it)
    numops, numiter = 1, 1
    tol = alg.tol
    HH = np.zeros((krylovdim + 1, krylovdim))
    converged = 0
    D = U = f = None
    while True:
        beta = fact.normres()
        K = len(fact)
        if beta <= tol and K < howmany and alg.verbosity >= WARN_LEVEL:
            warnings.warn(f"Invariant subspace of dimension {K} (up to requested tolerance `tol = {tol}`), "
                          f"which is smaller than the number of requested eigenvalues (i.e. `howmany == {howmany}`).")
        if K == krylovdim or beta <= tol or (alg.eager and K >= howmany):
            _t0 = _now()
            if K == 1:
                D = np.array([fact.alphas[0]])
                U = np.ones((1, 1))
                f = np.array([beta])
                converged = int(beta <= tol)
            else:
                dv, ev = fact.rayleighquotient()
                D, U = tridiageigh(dv, ev)
                p = eigsort(which)(D)
                D, U = permuteeig(D, U, p)
                f = U[K - 1, :] * beta
                converged = 0
                while converged < K and abs(f[converged]) <= tol:
                    converged += 1
            _tick("projected_eigenproblem", _t0)
            if converged >= howmany or beta <= tol:
                break
        if K < krylovdim:
            if alg.eager:
                fact = lz.expand_(it, fact)
                numops += 1
            else:
                # nothing happens between expansions until K == krylovdim or β <= tol
                # (eigsolve/lanczos.jl:45,77-79): run them back to back
                _t0 = _now()
                numops += lz.expand_many_(it, fact, krylovdim - K, tol)
                _tick("expand", _t0)
        else:
            if numiter == maxiter:
                break
            _t0 = _now()
            keep = (3 * krylovdim + 2 * converged) // 5
            # restore Lanczos form in the first keep columns — eigsolve/lanczos.jl:88-105
            if USE_NATIVE_RESTART:
                U = np.asfortranarray(U)
                Dc, fc = np.ascontiguousarray(D, dtype=np.float64), np.ascontiguousarray(f, dtype=np.float64)
                na, nb = np.empty(keep), np.empty(keep)
                dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
                L.check(L.load().b2k_host_lanczos_restart(K, keep, dp(Dc), dp(fc), dp(U), U.shape[0], dp(na), dp(nb)))
                fact.alphas[:keep] = na.tolist()
                fact.betas[:keep] = nb.tolist()
            else:
                restart_lanczos_form(HH, D, f, U, keep, fact.alphas, fact.betas)
            B = fact.basis()
            basistransform_(B, U[:, :keep])
            r = fact.residual()
            B[keep] = B[keep].scale_(1 / beta, r)     # B[keep+1] = scale!!(r, 1/β): column reuse
            fact = lz.shrink_(fact, keep)
            numiter += 1
            _tick("restart", _t0)
    _t0 = _now()
    hm = howmany
    if converged > howmany:
        hm = converged
    elif len(D) < howmany:
        hm = len(D)
    values = D[:hm].copy()
    B = fact.basis()
    if sink is None:
        vectors = _ritz_vectors(B, U, hm)
        r = fact.residual()
        residuals = [r.scale(U[-1, i]) for i in range(hm)]
    else:       # host-buffer path: stream the Ritz vectors out one at a time
        vectors, residuals = None, None
        for i in range(hm):
            v = B * U[:, i]
            sink(i, v)
            v.free()
    normres = np.abs(f[:hm])
    _tick("ritz_vectors", _t0)
    if converged < howmany and alg.verbosity >= WARN_LEVEL:
        warnings.warn(f"Lanczos eigsolve stopped without convergence after {numiter} iterations: "
                      f"{converged} eigenvalues converged, normres = {normres}, numops = {numops}")
    return values, vectors, ConvergenceInfo(converged, residuals, normres, numiter, numops)


def _eigsolve_blocklanczos(A, x0: "blz.Block", howmany: int, which: str, alg: BlockLanczos):
    """eigsolve(A, x₀::Block, howmany, which, alg::BlockLanczos) — src/eigsolve/blocklanczos.jl:1-144
    (SURVEY §8f-3).  Resolves degenerate eigenvalues up to the block size; thick restart as in the
    Lanczos driver, with `bs` Householder-restored rows instead of one."""
    maxiter, krylovdim = alg.maxiter, alg.krylovdim
    if howmany > krylovdim:
        raise ValueError(f"krylov dimension {krylovdim} too small to compute {howmany} eigenvalues")
    tol = alg.tol
    bs = len(x0)
    it = blz.BlockLanczosIterator(A, x0, krylovdim + bs, alg.orth, alg.qr_tol, fast_block=alg.fast_block)
    fact = blz.initialize(it)
    if alg.verbosity >= WARN_LEVEL and blz.warn_nonhermitian(fact.H[:fact.k, :fact.k]):
        warnings.warn("ignoring the antihermitian part of the block triangular matrix: "
                      "operator might not be hermitian?")
    numops, numiter, converged = bs + 1, 1, 0
    D = U = normresiduals = None
    while True:
        K = len(fact)
        beta = fact.normres()
        if beta < tol and K < howmany and alg.verbosity >= WARN_LEVEL:
            warnings.warn(f"Invariant subspace of dimension {K} (up to requested tolerance `tol = {tol}`), "
                          f"which is smaller than the number of requested eigenvalues (i.e. `howmany == {howmany}`).")
        if K >= krylovdim or beta <= tol or (alg.eager and K >= howmany):
            BTD = fact.H[:K, :K]
            D, U = np.linalg.eigh((BTD + BTD.T) / 2)           # eigen(Hermitian(BTD))
            D, U = permuteeig(D, U, eigsort(which)(D))
            bs_R = fact.R_size
            r = fact.residual()
            UU = U[K - bs_R:K, :]
            RR = blz.block_inner(r, r)
            normresiduals = np.sqrt(np.maximum(np.einsum("ik,ij,jk->k", UU, RR, UU), 0.0))
            converged = 0
            while converged < K and normresiduals[converged] <= tol:
                converged += 1
            if converged >= howmany or beta <= tol:
                break
        if K < krylovdim:
            blz.expand_(it, fact)
            numops += fact.R_size
        else:
            if numiter >= maxiter:
                break
            keep = max((3 * krylovdim + 2 * converged) // (5 * bs), 1) * bs
            H = np.zeros((keep + bs, keep))
            for j in range(keep):
                H[j, j] = D[j]
                H[keep:, j] = U[K - bs:K, j]
            for j in range(keep - 1, -1, -1):
                h, nu = householder_row(H, j + bs, range(0, j + 1), j)
                H[j + bs, j] = nu
                H[j + bs, :j] = 0
                lmul_householder(h, H)
                rmul_householder(H, h, slice(0, j + bs))
                rmul_householder(U, h)
            fact.H[:] = 0
            Hk = H[:keep, :keep]
            fact.H[:keep, :keep] = (Hk + Hk.T) / 2             # exactly symmetric
            B = fact.basis()
            basistransform_(B, U[:, :keep])
            R_new = OrthonormalBasis(fact.R.vec[:bs_R])
            basistransform_(R_new, H[keep + bs - bs_R:keep + bs, keep - bs_R:keep])
            fact.R.vec[:bs_R] = R_new.basis[:bs_R]
            fact.gram_R = None
            while len(B) > keep:
                B.pop().free()
            fact.k = keep
            numiter += 1
    hm = howmany
    if converged > howmany:
        hm = converged
    elif len(D) < howmany:
        hm = len(D)
    values = D[:hm].copy()
    V = fact.basis()
    vectors = [V * U[:, i] for i in range(hm)]
    bs_R, K = fact.R_size, len(fact)
    U2 = U[K - bs_R:K, :hm]
    Rb = OrthonormalBasis(fact.R.vec[:bs_R])
    residuals = [Rb * U2[:, i] for i in range(hm)]
    normresiduals = normresiduals[:hm]
    if converged < howmany and alg.verbosity >= WARN_LEVEL:
        warnings.warn(f"BlockLanczos eigsolve stopped without full convergence after {numiter} iterations: "
                      f"{converged} eigenvalues converged, normres = {normresiduals}, numops = {numops}")
    return values, vectors, ConvergenceInfo(converged, residuals, normresiduals, numiter, numops)
