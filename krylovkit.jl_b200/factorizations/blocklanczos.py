"""Block Lanczos factorization — mirror of src/factorizations/blocklanczos.jl.

A `Block` is a list of device vectors; the four block primitives run in libb200krylov
(b2k_block_inner / _axpy / _reorthogonalize / _qr), the bookkeeping of the block tridiagonal
matrix stays on the host as in the reference.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from .. import _lib as L
from ..algorithms import ModifiedGramSchmidt2, Orthogonalizer, mgs2
from ..operators import apply
from ..orthonormal import OrthonormalBasis
from ..vectors import B200Vec, handles


class Block:
    """Block(vec) — blocklanczos.jl:10-17: a non-empty list of vectors."""

    def __init__(self, vecs):
        vecs = list(vecs)
        if len(vecs) == 0:
            raise ValueError("blocklength must be >(0)")
        self.vec: list[B200Vec] = vecs

    def __len__(self):
        return len(self.vec)

    def __iter__(self):
        return iter(self.vec)

    def __getitem__(self, i):
        if isinstance(i, (int, np.integer)):
            return self.vec[i]
        if isinstance(i, slice):
            return Block(self.vec[i])
        return Block([self.vec[j] for j in i])

    def __setitem__(self, i, v):
        self.vec[i] = v

    def copy(self) -> "Block":
        return Block([v.copy() for v in self.vec])

    def norm(self) -> float:
        """norm(b::Block) = norm of the stacked vector — blocklanczos.jl:36."""
        return math.sqrt(sum(v.inner(v) for v in self.vec))

    @property
    def ctx(self):
        return self.vec[0].ctx


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def block_inner(B1: Block, B2: Block) -> np.ndarray:
    """block_inner — blocklanczos.jl:43-52: M[i, j] = <B1[i], B2[j]>, one device pass."""
    ctx = B1.ctx
    M = np.zeros((len(B1), len(B2)), order="F")
    ctx.check(ctx.lib.b2k_block_inner(ctx.h, handles(B1.vec), len(B1), handles(B2.vec), len(B2), _dptr(M)))
    return M


def block_axpy_(Y: Block, X, M: np.ndarray) -> Block:
    """Y[j] ← Y[j] − Σ_i X[i] M[i, j] — the double loops of blocklanczos.jl:177-181, 245-252."""
    ctx = Y.ctx
    X = list(X)
    M = np.asfortranarray(M, dtype=np.float64)
    ctx.check(ctx.lib.b2k_block_axpy(ctx.h, handles(Y.vec), len(Y), handles(X), len(X), _dptr(M), M.shape[0]))
    return Y


def block_reorthogonalize_(R: Block, V: OrthonormalBasis) -> Block:
    """block_reorthogonalize! — blocklanczos.jl:277-284."""
    ctx = R.ctx
    ctx.check(ctx.lib.b2k_block_reorthogonalize(ctx.h, handles(R.vec), len(R), handles(V.basis), len(V)))
    return R


def block_qr_(block: Block, tol: float):
    """block_qr! — blocklanczos.jl:312-353.  Returns (R[good, :], good_idx, is_drift)."""
    ctx = block.ctx
    p = len(block)
    R = np.zeros((p, p), order="F")
    good = (C.c_int32 * p)()
    drift = C.c_int32()
    ctx.check(ctx.lib.b2k_block_qr(ctx.h, handles(block.vec), p, float(tol), _dptr(R), good, C.byref(drift)))
    idx = [i for i in range(p) if good[i]]
    return R[idx, :], idx, bool(drift.value)


class BlockLanczosFactorization:
    """{k, V, H, R, R_size, norm_R} — blocklanczos.jl:82-94;  A V = V H + R Bᵀ with B = [0; I]."""

    def __init__(self, k, V: OrthonormalBasis, H: np.ndarray, R: Block, R_size: int, norm_R: float):
        self.k, self.V, self.H, self.R, self.R_size, self.norm_R = k, V, H, R, R_size, norm_R
        self.gram_R = None          # R'R when the flagged fast mode has it (saves CholeskyQR's first Gram pass)

    def __len__(self):
        return self.k

    def normres(self):
        return self.norm_R

    def basis(self):
        return self.V

    def residual(self) -> Block:
        return self.R[:self.R_size]


class BlockLanczosIterator:
    """BlockLanczosIterator(f, x₀, maxdim, orth, qr_tol) — blocklanczos.jl:131-157."""

    def __init__(self, operator, x0: Block, maxdim: int, orth: Orthogonalizer = mgs2, qr_tol: float = 1e-12,
                 fast_block: bool = False):
        self.fast_block = fast_block
        if x0.norm() < qr_tol:
            raise ValueError("initial vector should not have norm zero")
        if not isinstance(orth, ModifiedGramSchmidt2) and orth.tag != L.MGS2:
            raise ValueError("BlockLanczosIterator only supports ModifiedGramSchmidt2 orthogonalizer")
        self.operator, self.x0, self.maxdim, self.orth, self.qr_tol = operator, x0, maxdim, orth, qr_tol


def warn_nonhermitian(M: np.ndarray) -> bool:
    """blocklanczos.jl:286-291 — True when M is not symmetric to eps^(2/5)."""
    return not np.allclose(M, M.T, rtol=math.sqrt(np.finfo(np.float64).eps),
                           atol=np.finfo(np.float64).eps ** 0.4)


# FLAGGED B200-first mode (off = the reference's arithmetic: modified Gram-Schmidt everywhere).  When on,
# block_reorthogonalize! runs as block classical Gram-Schmidt twice (the basis is read once per pass for the whole
# block instead of once per vector and basis column) and block_qr! as CholeskyQR2; the QR factor is the same matrix
# (the QR factorization with positive diagonal is unique), the orthogonalisation coefficients differ at rounding
# level, Ritz values agree with the reference mode to ~1e-12.  Rank-deficient blocks fall back to the reference's
# block_qr!.  kk.BlockLanczos(..., fast_block=True) switches it on per solve.
FAST_BLOCK = False


def _apply_block(operator, X: Block) -> Block:
    """apply(f, block::Block) — blocklanczos.jl:38.  A device CSR operator reads the matrix once for the whole
    block (b2k_op_apply_block, SpMM); same bits as the loop of single applies."""
    from ..operators import B200CSR
    if isinstance(operator, B200CSR) and len(X) > 1:
        ctx = X.ctx
        Y = [ctx.empty(x.space) for x in X]
        ctx.check(ctx.lib.b2k_op_apply_block(ctx.h, operator.h, handles(X.vec), handles(Y), len(X)))
        return Block(Y)
    return Block([apply(operator, x) for x in X])


def block_orthogonalize_fast_(R: Block, V: OrthonormalBasis, want_gram: bool = True):
    """flagged: BCGS2 of the block R against V in 4 sweeps over V; returns (V'R summed over the passes, R'R)."""
    ctx = R.ctx
    p, k = len(R), len(V)
    H = np.zeros((k, p), order="F")
    G = np.zeros((p, p), order="F")
    ctx.check(ctx.lib.b2k_block_orthogonalize(ctx.h, handles(R.vec), p, handles(V.basis), k, 2, _dptr(H),
                                              _dptr(G) if want_gram else None))
    return H, G


def block_cholqr_(block: Block, tol: float, G0: np.ndarray | None = None):
    """flagged: CholeskyQR2 in place.  Returns (R, ok); ok = False: the block is numerically rank deficient at the
    scale block_qr! drops vectors — nothing was changed, use block_qr_ (reference MGS with rank detection)."""
    ctx = block.ctx
    p = len(block)
    R = np.zeros((p, p), order="F")
    ok = C.c_int32()
    g0 = _dptr(np.asfortranarray(G0, dtype=np.float64)) if G0 is not None else None
    ctx.check(ctx.lib.b2k_block_cholqr(ctx.h, handles(block.vec), p, float(tol), g0, _dptr(R), C.byref(ok)))
    return R, bool(ok.value)


def initialize(it: BlockLanczosIterator) -> BlockLanczosFactorization:
    """initialize(iter) — blocklanczos.jl:159-190.  (The reference applies the operator once to
    x₀[1] only to fix a number type; the application is kept because eigsolve counts it.)"""
    X0 = it.x0
    if X0.norm() == 0:
        raise ValueError("initial vector should not have norm zero")
    apply(it.operator, X0[0]).free()
    X1 = X0.copy()
    _, good, _ = block_qr_(X1, it.qr_tol)
    X1 = X1[good]
    V = OrthonormalBasis(X1.vec)
    bs = len(X1)
    AX1 = _apply_block(it.operator, X1)
    M1 = block_inner(X1, AX1)
    BTD = np.zeros((it.maxdim, it.maxdim))
    BTD[:bs, :bs] = M1
    block_axpy_(AX1, X1.vec, M1)
    return BlockLanczosFactorization(bs, V, BTD, AX1, bs, AX1.norm())


def block_lanczosrecurrence(operator, V: OrthonormalBasis, B: np.ndarray, fast: bool = False):
    """block_lanczosrecurrence(…, ::ModifiedGramSchmidt2) — blocklanczos.jl:232-251.
    fast (flagged): AX is orthogonalised against ALL of V by BCGS2 — which removes the X M and Xprev B' terms of
    the three-term recurrence as part of the same sweeps (X, Xprev are columns of V) — and M = X'AX is read off the
    summed coefficients, like lanczos.jl:313-324 does for the single-vector CGS2 recurrence.  Returns the Gram
    matrix of the new residual as a third value (None in the reference mode)."""
    bs, bs_prev = B.shape
    k = len(V)
    X = Block(V[k - bs:k])
    AX = _apply_block(operator, X)
    if fast and bs <= 8:
        H, G = block_orthogonalize_fast_(AX, V)
        return AX, H[k - bs:k, :].copy(), G
    M = block_inner(X, AX)
    Xprev = V[k - bs_prev - bs:k - bs]
    # AX[j] -= Σ_i X[i] M[i,j] + Σ_i Xprev[i] B[j,i]: one fused sweep with the stacked coefficients
    block_axpy_(AX, X.vec + list(Xprev), np.vstack([M, B.T[:len(Xprev), :]]))
    block_reorthogonalize_(AX, V)
    return AX, M, None


def expand_(it: BlockLanczosIterator, state: BlockLanczosFactorization) -> BlockLanczosFactorization:
    """expand!(iter, state) — blocklanczos.jl:192-230."""
    k = state.k
    R = state.R[:state.R_size]
    bs = len(R)
    V = state.V
    fast = (FAST_BLOCK or it.fast_block) and bs <= 8
    B = good = None
    if fast:
        B, ok = block_cholqr_(R, it.qr_tol, state.gram_R)
        if ok:
            good = list(range(bs))
        else:
            fast = False                               # rank-deficient block: this step runs in the reference mode
    if good is None:
        Rcopy = R.copy()
        B, good, drift = block_qr_(R, it.qr_tol)
        if drift:
            # an excessively small β in block_qr! lets the column space of R drift: re-project and redo
            block_reorthogonalize_(R, V)
            _, good, drift = block_qr_(R, it.qr_tol)
            B = block_inner(R[good], Rcopy)            # keeps R = X B
    bs_next = len(good)
    for i in good:
        V.push(R[i])
    state.H[k:k + bs_next, k - bs:k] = B[:bs_next, :bs]
    state.H[k - bs:k, k:k + bs_next] = B[:bs_next, :bs].T
    Rnext, Mnext, gram = block_lanczosrecurrence(it.operator, V, B, fast)
    state.H[k:k + bs_next, k:k + bs_next] = Mnext[:bs_next, :bs_next]
    state.R.vec[:bs_next] = Rnext.vec
    state.gram_R = gram
    state.norm_R = Rnext.norm() if gram is None else math.sqrt(max(float(np.trace(gram)), 0.0))
    state.k += bs_next
    state.R_size = bs_next
    return state
