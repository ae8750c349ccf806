"""The flagged one-pass mode of the Golub-Kahan-Lanczos step: `b2k_op_apply_normal_gram` — y = A x and
z = A'(A x) from ONE pass over a dense A — and `GKL(onepass=True)` on top of it (factorizations/gkl.py).

The kernel against the two separate products (and a Float64 numpy product), the solver against the oracle's two-pass
svdsolve and the Float64 truth.  The bodies also run on the numpy stand-in (tests/test_hostsim.py).

(The file sorts last on purpose: the kernel of this mode was written after the round's GPU budget had ended and
has not run on a B200 yet; under `pytest -x` everything else is through before it is reached.)"""
import numpy as np
import pytest

import krylovkit_jl_b200 as kk
from krylovkit_jl_b200.operators import apply_adjoint, apply_normal, apply_normal_gram
from oracle import krylov_oracle as ko

pytestmark = pytest.mark.gpu
SEED = 20260923


@pytest.fixture()
def variant_b():
    """kernel variant B (64-row tiles, register-pipelined; Float32 with n <= 512, everything else runs variant A)"""
    lib = kk._lib.load()
    assert lib.b2k_debug_set_onepass_variant(1) == 0
    yield
    assert lib.b2k_debug_set_onepass_variant(0) == 0


@pytest.mark.parametrize("m,n", [(20000, 512), (33, 300), (70001, 256), (96, 500), (160, 512), (100, 70), (4000, 6)])
def test_apply_normal_gram_variant_b(variant_b, m, n):
    test_apply_normal_gram(m, n, np.float32)


def test_svdsolve_onepass_config4_small_f32_variant_b(variant_b):
    test_svdsolve_onepass_config4_small_f32()


@pytest.mark.parametrize("m,n,dtype", [(20000, 512, np.float32), (1000, 70, np.float64), (33, 300, np.float32),
                                       (5000, 1030, np.float32), (4097, 600, np.float64), (96, 6, np.float64),
                                       (70001, 256, np.float32)])
def test_apply_normal_gram(m, n, dtype):
    """y bit-for-bit a product A x up to summation order (checked against Float64 numpy), z = A'(A x); x untouched;
    two calls give identical bits (fixed reduction order)."""
    rng = np.random.default_rng(m + n)
    A = (rng.random((m, n)) - 0.5).astype(dtype)
    x = (rng.random(n) - 0.5).astype(dtype)
    ctx = kk.B200Context(m, 6, dtype=dtype)
    sv = ctx.add_space(n, 8, sharded=False)
    op = kk.B200Dense.from_host(ctx, A, sv)
    xv = ctx.from_host(x, space=sv)
    y, z = apply_normal_gram(op, xv)
    A64 = A.astype(np.float64)
    y64 = A64 @ x.astype(np.float64)
    eps = np.finfo(dtype).eps
    yh, zh = y.to_host(), z.to_host()
    assert np.abs(yh - y64).max() <= 8 * eps * np.sqrt(n) * np.abs(y64).max() + 1e-300
    z64 = A64.T @ yh.astype(np.float64)              # A' applied to the y that was actually formed
    assert np.abs(zh - z64).max() <= 8 * eps * np.sqrt(m) * np.abs(z64).max() + 1e-300
    np.testing.assert_array_equal(xv.to_host(), x)
    if n <= 512:          # the separate products of the library (config 4's width at most) agree to the same accuracy
        y2 = apply_normal(op, xv)
        z2 = apply_adjoint(op, y2)
        assert np.abs(y2.to_host() - yh).max() <= 8 * eps * np.sqrt(n) * np.abs(y64).max()
        assert np.abs(z2.to_host().astype(np.float64) - zh).max() <= 16 * eps * np.sqrt(m) * np.abs(z64).max()
        del y2, z2
    y3, z3 = apply_normal_gram(op, xv)
    np.testing.assert_array_equal(y3.to_host(), yh)
    np.testing.assert_array_equal(z3.to_host(), zh)
    ctx.close()


def test_apply_normal_gram_errors():
    ctx = kk.B200Context(200, 6)
    sv = ctx.add_space(10, 6, sharded=False)
    op = kk.B200Dense.from_host(ctx, np.ones((200, 10)), sv)
    x = ctx.from_host(np.ones(10), space=sv)
    y, y2, y3, z = ctx.empty(0), ctx.empty(0), ctx.empty(0), ctx.empty(sv)
    with pytest.raises(ValueError):                   # z aliases x
        ctx.check(ctx.lib.b2k_op_apply_normal_gram(ctx.h, op.h, x.handle, y.handle, x.handle))
    with pytest.raises(kk.DimensionMismatch):         # y in the wrong space
        ctx.check(ctx.lib.b2k_op_apply_normal_gram(ctx.h, op.h, x.handle, x.handle, z.handle))
    csr = kk.B200CSR.stencil(ctx, 20, 10)
    with pytest.raises(kk.B200Error):                 # dense operators only
        ctx.check(ctx.lib.b2k_op_apply_normal_gram(ctx.h, csr.h, y.handle, y2.handle, y3.handle))
    ctx.close()


@pytest.mark.parametrize("orth,oorth", [(kk.mgs2, ko.Orth(ko.MGS2)), (kk.cgs2, ko.Orth(ko.CGS2)),
                                        (kk.ClassicalGramSchmidtIR(eta=0.75), ko.Orth(ko.CGSIR, 0.75))],
                         ids=["mgs2", "cgs2", "cgsr"])
def test_svdsolve_onepass_f64(orth, oorth):
    """Float64, a spread spectrum (alpha_k > beta_k after the first restart: the recycling recursion would amplify its
    errors ~5x per step — the error estimate makes those steps fall back to a direct A'u): numops / numiter of the
    oracle's two-pass svdsolve, singular values to 1e-10, the triplets satisfy the SVD relations, fewer passes."""
    m, n = 3001, 120
    rng = np.random.default_rng(11)
    A = rng.standard_normal((m, n))
    u0 = rng.standard_normal(m)
    alg = kk.GKL(orth=orth, krylovdim=25, maxiter=100, tol=1e-10, verbosity=0, onepass=True)
    S, Lv, Rv, info = kk.svdsolve(A, u0, 5, "LR", alg)
    oS, _, _, oinfo = ko.svdsolve_gkl(A, u0, 5, "LR", krylovdim=25, maxiter=100, tol=1e-10, orth=oorth)
    ref = np.linalg.svd(A, compute_uv=False)
    assert info.converged >= 5
    # (identical on the numpy stand-in; on the device a convergence check may flip at the threshold)
    assert abs(info.numiter - oinfo["numiter"]) <= 1 and abs(info.numops - oinfo["numops"]) <= 25
    np.testing.assert_allclose(S[:5], ref[:5], rtol=1e-10)
    np.testing.assert_allclose(S[:5], oS[:5], rtol=1e-10)
    U, V = np.column_stack(Lv), np.column_stack(Rv)
    c = U.shape[1]
    assert np.abs(U.T @ U - np.eye(c)).max() < 1e-9
    np.testing.assert_allclose(A.T @ U, V * S[:c], atol=1e-7)
    Rm = np.column_stack(info.residual)
    np.testing.assert_allclose(A @ V, U * S[:c] + Rm, atol=1e-7)
    assert info.numops // 2 + 1 <= info.passes <= (3 * info.numops) // 4       # two-pass step: passes == numops


def test_svdsolve_onepass_config4_small_f32():
    """config 4 at test size in the one-pass mode (dense tall Float32, 6 triplets, GKL(krylovdim=30, tol=1e-5), the
    reference-default MGS2): converged values within the Float32 bar of the Float64 truth, fewer passes than the
    two-pass run for the same kind of work.  (At this test size the spectrum is 16 % wide — m / n = 39 — so part of the
    steps fall back to a direct A'u; at config 4's m / n = 3900 the spectrum is flat and every step recycles: bench.py's
    c4 record reports the passes of the full-size run.)"""
    m, n = 20000, 512
    A = ko.dense_splitmix(SEED, m, n)
    u0 = ko.splitmix_vector(SEED + 1, m, dtype=np.float32)
    ctx = kk.B200Context(m, 56, dtype=np.float32)
    sv = ctx.add_space(n, 84, sharded=False)
    op = kk.B200Dense.splitmix(ctx, m, n, SEED, sv)
    ref = np.linalg.svd(A.astype(np.float64), compute_uv=False)
    out = {}
    for onepass in (False, True):
        alg = kk.GKL(orth=kk.mgs2, krylovdim=30, maxiter=100, tol=1e-5, verbosity=0, onepass=onepass)
        S, Lv, Rv, info = kk.svdsolve(op, ctx.from_host(u0), 6, "LR", alg)
        assert info.converged >= 6
        np.testing.assert_allclose(S[:6], ref[:6], rtol=3e-5)
        u, v = Lv[0].to_host().astype(np.float64), Rv[0].to_host().astype(np.float64)
        assert np.linalg.norm(A.astype(np.float64) @ v - S[0] * u) < 1e-3 * S[0]
        assert np.linalg.norm(A.astype(np.float64).T @ u - S[0] * v) < 1e-3 * S[0]
        out[onepass] = (info.numops, info.numiter, info.passes)
        del Lv, Rv, info
    assert out[False][2] == out[False][0]                                   # the reference's step: two passes per step
    assert out[True][0] // 2 + 1 <= out[True][2] <= (4 * out[True][0]) // 5
    # two roundings of the same recurrence: the number of restart cycles may differ a little, not the kind of work
    assert abs(out[True][1] - out[False][1]) <= 2
    ctx.close()
