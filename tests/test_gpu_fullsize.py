"""Parity at the sizes BASELINE.json quotes (n = 1e6 .. 1e7): the CUDA path through the C-ABI against the ORACLE's
results on the same (A, x0), committed in tests/golden/fullsize.json by tests/golden/make_fullsize_golden.py (the
oracle needs minutes per case at this size; the GPU needs well under a second).

Tolerances: Ritz values 1e-10 relative (north_star) plus a floor of 4 ulp(||A||) — 1e-10 of lambda_1 ~ 1e-5 is
1.5 ulp of ||A|| = 8, below what two correct summation orders agree to; GMRES residual norms 1e-8 relative,
solution entries 1e-8; `numops` / `numiter` identical."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import krylovkit_jl_b200 as kk

SEED = 20260923
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "fullsize.json")))
EPS = float(np.finfo(np.float64).eps)


def _check_ritz(vals, info, g, norm_a):
    ref = np.array(g["ritz"])
    diff = np.abs(np.array(vals[:len(ref)]) - ref)
    assert np.all(diff <= 1e-10 * np.abs(ref) + 4 * norm_a * EPS), (list(vals[:len(ref)]), list(ref), list(diff / np.abs(ref)))
    assert info.numops == g["numops"] and info.numiter == g["numiter"]
    np.testing.assert_allclose(info.normres[:len(ref)], g["normres"], rtol=1e-6, atol=1e-13)


@pytest.mark.parametrize("case,cycles", [("c2", 2), ("c2_1e6", 5), ("c2_mgs2", 2), ("c5s", 5)])
def test_lanczos_eigsolve_full_size_matches_oracle(case, cycles):
    """configs[1] (1e7-row 5-point Laplacian, krylovdim 60, CGS2 and the reference-default MGS2), its 1e6-row
    sibling through all 5 restart cycles, and the configs[4] operator (7-point, krylovdim 30) at 8e6 rows."""
    c = GOLD[case]
    g = c["after_cycles"][str(cycles)]
    nx, ny, nz = c["grid"]
    n = nx * ny * nz
    ctx = kk.B200Context(n, c["krylovdim"] + 16)
    if nz > 1:
        op = kk.B200CSR.stencil(ctx, nx, ny, nz, (6.0, -1, -1, -1, -1, -1, -1))
        norm_a = 12.0
    else:
        op = kk.B200CSR.stencil(ctx, nx, ny)
        norm_a = 8.0
    orth = {"cgs2": kk.cgs2, "mgs2": kk.mgs2}[c["orth"]]
    alg = kk.Lanczos(orth=orth, krylovdim=c["krylovdim"], maxiter=cycles, tol=0.0, verbosity=0)
    vals, vecs, info = kk.eigsolve(op, ctx.splitmix(SEED), c["howmany"], "SR", alg)
    _check_ritz(vals, info, g, norm_a)
    # residual identity of the returned pairs on the device: ||A v - lambda v|| = normres (test/eigsolve.jl:74-82)
    w = kk.apply(op, vecs[0])
    w.add_(vecs[0], -float(vals[0]))
    assert abs(w.norm() - info.normres[0]) <= 1e-8 * max(info.normres[0], 1e-8)
    ctx.close()


@pytest.mark.parametrize("case", ["c3", "c3_mgs2"])
def test_gmres_full_size_matches_oracle(case):
    """configs[2]: restarted GMRES(40) on the 1e7-row convection-diffusion operator, b = A*1, 2 restart cycles."""
    c = GOLD[case]
    g = c["after_cycles"]["2"]
    nx, ny, _ = c["grid"]
    n = nx * ny
    ctx = kk.B200Context(n, c["krylovdim"] + 16)
    op = kk.B200CSR.stencil(ctx, nx, ny, 1, tuple(c["coeffs"]))
    ones = ctx.full(1.0)
    b = kk.apply(op, ones)
    orth = {"cgs2": kk.cgs2, "mgs2": kk.mgs2}[c["orth"]]
    alg = kk.GMRES(orth=orth, krylovdim=c["krylovdim"], maxiter=2, tol=1e-300, verbosity=0)
    x, info = kk.linsolve(op, b, None, alg)
    assert info.numops == g["numops"] and info.numiter == g["numiter"]
    assert abs(info.normres - g["normres"]) <= 1e-8 * g["normres"]
    assert abs(x.norm() - g["x_norm"]) <= 1e-10 * g["x_norm"]
    xh = x.to_host()
    for i, v in g["x_samples"].items():
        assert abs(xh[int(i)] - v) <= 1e-8 * max(abs(v), 1e-3), (i, xh[int(i)], v)
    # b = A x + r (test/linsolve.jl:230), evaluated on the device
    chk = kk.apply(op, x)
    chk.add_(info.residual, 1.0).add_(b, -1.0)
    assert chk.norm() <= 1e-10 * b.norm()
    ctx.close()


@pytest.mark.parametrize("orth_name", ["mgs2", "cgsr"])
def test_svdsolve_config4_full_size_vs_float64_truth(orth_name):
    """configs[3] at its full size (dense 2e6 x 512 Float32 splitmix matrix, GKL krylovdim 30, tol 1e-5, 6 triplets :LR)
    through the reference's two-pass step, with the reference-default MGS2 and with the iterative-refinement
    orthogonalizer the reference's own Float32 tests use: the converged singular values against the Float64 truth of the
    same matrix (tests/golden/fullsize.json:c4_truth — sqrt eig of the Float64 Gram matrix) within the Float32 bar 3e-5,
    and the SVD relations A v = s u (+ residual), A'u = s v evaluated on the device."""
    g = GOLD["c4_truth"]
    m, n = g["shape"]
    kd = 30
    ctx = kk.B200Context(m, kd + 24, dtype=np.float32)
    sv = ctx.add_space(n, kd + 24, sharded=False)
    op = kk.B200Dense.splitmix(ctx, m, n, g["seed"], sv)
    u0 = ctx.splitmix(g["seed"] + 1)
    orth = {"mgs2": kk.mgs2, "cgsr": kk.ClassicalGramSchmidtIR(eta=0.75)}[orth_name]
    alg = kk.GKL(orth=orth, krylovdim=kd, maxiter=100, tol=1e-5, verbosity=0)
    S, Lv, Rv, info = kk.svdsolve(op, u0, 6, "LR", alg)
    assert info.converged >= 6
    ref = np.array(g["sigma_float64_truth"])
    rel = np.abs(np.array(S[:6], dtype=np.float64) - ref) / ref
    assert rel.max() <= 3e-5, (list(S[:6]), list(ref))
    for i in range(3):
        w = kk.apply_normal(op, Rv[i])
        w.add_(Lv[i], -float(S[i]))
        z = kk.apply_adjoint(op, Lv[i])
        z.add_(Rv[i], -float(S[i]))
        assert w.norm() / S[i] < 1e-4 and z.norm() / S[i] < 1e-4
        del w, z
    assert info.passes == info.numops           # the reference's step: both products stream A
    ctx.close()
