"""CPU tests that pin the oracle (oracle/krylov_oracle.py) against the reference's own
known-answer fixtures and against the invariants the reference asserts in its test-suite
(test/factorize.jl, test/linalg.jl, test/eigsolve.jl, test/linsolve.jl, test/svdsolve.jl,
test/issues.jl, test/block.jl)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import krylov_oracle as ko

HERE = os.path.dirname(os.path.abspath(__file__))
EPS = np.finfo(np.float64).eps
TOL = EPS ** (2 / 3)              # tolerance(T) — test/testsetup.jl:14
ORTHS = {
    "cgs": ko.Orth(ko.CGS), "mgs": ko.Orth(ko.MGS), "cgs2": ko.Orth(ko.CGS2), "mgs2": ko.Orth(ko.MGS2),
    "cgsr": ko.Orth(ko.CGSIR, 0.75), "mgsr": ko.Orth(ko.MGSIR, 0.75),      # test/runtests.jl:18-24
}
REORTH = ["cgs2", "mgs2", "cgsr", "mgsr"]     # Lanczos/GKL tests skip plain cgs/mgs (factorize.jl:7)


def symm(rng, n):
    A = rng.standard_normal((n, n))
    return (A + A.T) / 2


# ---- known-answer fixtures of the reference ---------------------------------------------

def test_issue143_matrix_spectrum():
    """test/issues.jl:39-129: the literal 71x71 matrix; a complete Lanczos factorization
    must reproduce the dense spectrum."""
    A = np.load(os.path.join(HERE, "golden", "issue143_matrix.npy"))
    n = A.shape[0]
    ref = np.linalg.eigvalsh(A)
    rng = np.random.default_rng(143)
    D, V, info = ko.eigsolve_lanczos(A, rng.standard_normal(n), n, "SR", krylovdim=n, maxiter=1,
                                     tol=1e-12, orth=ORTHS["mgs2"])
    assert len(D) == n
    np.testing.assert_allclose(D, ref, rtol=1e-10, atol=1e-10 * np.abs(ref).max())
    U = np.column_stack(V)
    np.testing.assert_allclose(A @ U, U * D, atol=1e-8 * np.abs(ref).max())


def test_issue156_identity():
    """test/issues.jl:32-36: eigsolve([1 0; 0 1]) -> 1.0."""
    A = np.eye(2)
    D, V, info = ko.eigsolve_lanczos(A, np.array([0.3, 0.8]), 1, "LM", krylovdim=2, tol=1e-12)
    assert info["converged"] >= 1
    assert np.allclose(D[: info["converged"]], 1.0)


def test_toric_code_ground_energy():
    """test/eigsolve.jl:471-549: -H of the 3x3 toric code has lowest eigenvalue -16."""
    H = ko.toric_code_hamiltonian(3, 3)
    assert H.shape == (2 ** 18, 2 ** 18)
    assert abs(H - H.T).max() == 0
    x0 = ko.splitmix_vector(7, H.shape[0])
    D, V, info = ko.eigsolve_lanczos(-H, x0, 1, "SR", krylovdim=30, maxiter=30, tol=1e-8,
                                     orth=ORTHS["cgs2"])
    assert info["converged"] >= 1
    assert abs(D[0] + 16.0) < 1e-8


def test_laplacian_closed_form_config1():
    """BASELINE config 1: eigsolve(Lanczos, :SR, 4) on the 1e4 x 1e4 5-point Laplacian,
    krylovdim = 30, against the closed-form spectrum.  A non-degenerate
    125 x 80 grid is used so that the four smallest eigenvalues are simple."""
    nx, ny = 125, 80
    A = ko.stencil_matrix(nx, ny)
    lam = ko.laplace_eigenvalues(nx, ny)
    x0 = ko.splitmix_vector(20260923, nx * ny)
    D, V, info = ko.eigsolve_lanczos(A, x0, 4, "SR", krylovdim=30, maxiter=300, tol=1e-10,
                                     orth=ORTHS["mgs2"])
    assert info["converged"] >= 4
    np.testing.assert_allclose(D[:4], lam[:4], rtol=1e-10)
    for i in range(4):
        assert np.linalg.norm(A @ V[i] - D[i] * V[i]) < 1e-8


# ---- invariants of the reference's tests -------------------------------------------------

@pytest.mark.parametrize("name", list(ORTHS))
def test_orthonormalize_identities(name):
    """test/linalg.jl:4-25."""
    rng = np.random.default_rng(1)
    n, k = 100, 10
    Q, _ = np.linalg.qr(rng.standard_normal((n, k)))
    b = [Q[:, j].copy() for j in range(k)]
    a = rng.standard_normal(n)
    v, beta, x = ko.orthonormalize(a.copy(), b, np.zeros(k), ORTHS[name])
    assert abs(np.linalg.norm(v) - 1) < 1e-12
    assert abs(np.hypot(beta, np.linalg.norm(x)) - np.linalg.norm(a)) < 1e-10
    assert np.abs(Q.T @ v).max() < (1e-8 if name in ("cgs", "mgs") else 1e-12)
    np.testing.assert_allclose(Q @ x + beta * v, a, atol=1e-10)


@pytest.mark.parametrize("name", REORTH)
def test_lanczos_factorization_invariants(name):
    """test/factorize.jl:140-158: after every expand!: V'V = I, ‖r‖ = β, A V = V H + r e';
    again after shrink!."""
    rng = np.random.default_rng(2)
    n = 100
    A = symm(rng, n)
    f = ko.lanczos_initialize(A, rng.standard_normal(n), ORTHS[name])
    for _ in range(29):
        f = ko.lanczos_expand(A, f, ORTHS[name])
        V = np.column_stack(f.V)
        k = f.k
        T = np.diag(f.alphas) + np.diag(f.betas[:k - 1], 1) + np.diag(f.betas[:k - 1], -1)
        assert np.abs(V.T @ V - np.eye(k)).max() < 1e-12
        assert abs(np.linalg.norm(f.r) - f.normres()) < 1e-12
        E = A @ V - V @ T
        E[:, -1] -= f.r
        assert np.abs(E).max() < 1e-11
    f = ko.lanczos_shrink(f, 12)
    V = np.column_stack(f.V)
    T = np.diag(f.alphas) + np.diag(f.betas[:11], 1) + np.diag(f.betas[:11], -1)
    E = A @ V - V @ T
    E[:, -1] -= f.r
    assert f.k == 12 and len(f.V) == 12 and np.abs(E).max() < 1e-11


@pytest.mark.parametrize("name", list(ORTHS))
def test_arnoldi_factorization_invariants(name):
    """test/factorize.jl:185-203."""
    rng = np.random.default_rng(3)
    n = 100
    A = rng.standard_normal((n, n))
    f = ko.arnoldi_initialize(A, rng.standard_normal(n), ORTHS[name])
    tol = 1e-9 if name in ("cgs", "mgs") else 1e-11
    for _ in range(25):
        f = ko.arnoldi_expand(A, f, ORTHS[name])
    k = f.k
    V = np.column_stack(f.V)
    H = np.zeros((k, k))
    for j in range(1, k + 1):
        for i in range(1, min(j + 1, k) + 1):
            H[i - 1, j - 1] = f.h(i, j)
    assert np.abs(V.T @ V - np.eye(k)).max() < tol
    E = A @ V - V @ H
    E[:, -1] -= f.r
    assert np.abs(E).max() < tol
    assert abs(np.linalg.norm(f.r) - f.normres()) < 1e-12


@pytest.mark.parametrize("name", REORTH)
def test_gkl_factorization_invariants(name):
    """test/factorize.jl:285-309: U'U = I, V'V = I, A V = U B + r e', A'U = V B'."""
    rng = np.random.default_rng(4)
    m, n = 120, 80
    A = rng.standard_normal((m, n))
    f = ko.gkl_initialize(A, rng.standard_normal(m), ORTHS[name])
    for _ in range(20):
        f = ko.gkl_expand(A, f, ORTHS[name])
    k = f.k
    U, V = np.column_stack(f.U), np.column_stack(f.V)
    B = np.diag(f.alphas) + np.diag(f.betas[:k - 1], -1)
    assert np.abs(U.T @ U - np.eye(k)).max() < 1e-11
    assert np.abs(V.T @ V - np.eye(k)).max() < 1e-11
    E = A @ V - U @ B
    E[:, -1] -= f.r
    assert np.abs(E).max() < 1e-10
    assert np.abs(A.T @ U - V @ B.T).max() < 1e-10


@pytest.mark.parametrize("name", REORTH)
@pytest.mark.parametrize("which", ["SR", "LR", "LM"])
def test_eigsolve_iterative(name, which):
    """test/eigsolve.jl:84-136: values vs dense eigvals, U'U = I, A U = U D + R."""
    rng = np.random.default_rng(5)
    n = 100
    A = symm(rng, n)
    ref = np.linalg.eigvalsh(A)
    order = {"SR": ref, "LR": ref[::-1], "LM": ref[np.argsort(-np.abs(ref))]}[which]
    D, V, info = ko.eigsolve_lanczos(A, rng.standard_normal(n), 6, which, krylovdim=30, maxiter=50,
                                     tol=TOL, orth=ORTHS[name])
    c = info["converged"]
    assert c >= 6
    np.testing.assert_allclose(D[:c], order[:c], atol=10 * TOL * np.abs(ref).max())
    U = np.column_stack(V)
    assert np.abs(U.T @ U - np.eye(U.shape[1])).max() < 1e-10
    Rm = np.column_stack(info["residual"])
    np.testing.assert_allclose(A @ U, U * D[: U.shape[1]] + Rm, atol=1e-9)


@pytest.mark.parametrize("name", list(ORTHS))
def test_gmres(name):
    """test/linsolve.jl:119-285: b = A x; restarted: b = A x + residual; shifted operator."""
    rng = np.random.default_rng(6)
    n = 100
    A = rng.standard_normal((n, n)) / np.sqrt(n) + 3 * np.eye(n)
    b = rng.standard_normal(n)
    x, info = ko.linsolve_gmres(A, b, krylovdim=n, maxiter=2, tol=1e-12 * np.linalg.norm(b), orth=ORTHS[name])
    assert info["converged"] == 1
    np.testing.assert_allclose(A @ x, b, atol=1e-9)
    x, info = ko.linsolve_gmres(A, b, krylovdim=12, maxiter=3, tol=1e-14, orth=ORTHS[name])
    np.testing.assert_allclose(A @ x + info["residual"], b, atol=1e-9)
    x, info = ko.linsolve_gmres(A, b, krylovdim=20, maxiter=50, tol=1e-10, orth=ORTHS[name], a0=0.7, a1=-1.3)
    assert info["converged"] == 1
    np.testing.assert_allclose(0.7 * x - 1.3 * (A @ x), b, atol=1e-8)


@pytest.mark.parametrize("name", REORTH)
def test_svdsolve(name):
    """test/svdsolve.jl:14,60-63,88-98."""
    rng = np.random.default_rng(7)
    m, n = 150, 90
    A = rng.standard_normal((m, n))
    ref = np.linalg.svd(A, compute_uv=False)
    S, Lv, Rv, info = ko.svdsolve_gkl(A, rng.standard_normal(m), 5, "LR", krylovdim=25, maxiter=100,
                                      tol=1e-10, orth=ORTHS[name])
    assert info["converged"] >= 5
    c = len(S)
    np.testing.assert_allclose(S, ref[:c], rtol=1e-9)
    U, V = np.column_stack(Lv), np.column_stack(Rv)
    assert np.abs(U.T @ U - np.eye(c)).max() < 1e-9
    assert np.abs(V.T @ V - np.eye(c)).max() < 1e-9
    np.testing.assert_allclose(A.T @ U, V * S, atol=1e-8)
    Rm = np.column_stack(info["residual"])
    np.testing.assert_allclose(A @ V, U * S + Rm, atol=1e-8)


def test_svdsolve_float32():
    rng = np.random.default_rng(8)
    m, n = 400, 64
    A = (rng.random((m, n)) - 0.5).astype(np.float32)
    ref = np.linalg.svd(A.astype(np.float64), compute_uv=False)
    S, Lv, Rv, info = ko.svdsolve_gkl(A, rng.random(m).astype(np.float32), 6, "LR", krylovdim=30,
                                      maxiter=100, tol=1e-5, orth=ORTHS["cgs2"])
    assert info["converged"] >= 6
    np.testing.assert_allclose(S[:6], ref[:6], rtol=5e-5)


def test_block_primitives():
    """test/block.jl:74-183."""
    rng = np.random.default_rng(9)
    n, p = 200, 6
    X = [rng.standard_normal(n) for _ in range(p)]
    Y = [rng.standard_normal(n) for _ in range(p)]
    np.testing.assert_allclose(ko.block_inner(X, Y), np.column_stack(X).T @ np.column_stack(Y), atol=1e-12)
    Q, _ = np.linalg.qr(rng.standard_normal((n, 12)))
    V = [Q[:, j].copy() for j in range(12)]
    Rb = ko.block_reorthogonalize([x.copy() for x in X], V)
    assert np.abs(Q.T @ np.column_stack(Rb)).max() < 1e-12
    X2 = [x.copy() for x in X]
    X2[3] = 2 * X2[0] - X2[1]                    # rank deficient
    X0 = np.column_stack(X2)
    R, good, drift = ko.block_qr(X2, 1e-10)
    assert good == [0, 1, 2, 4, 5]
    Qg = np.column_stack([X2[i] for i in good])
    np.testing.assert_allclose(Qg @ R, X0, atol=1e-9)
    assert np.abs(Qg.T @ Qg - np.eye(5)).max() < 1e-12


def test_householder_givens_basis_equal_dense():
    """test/linalg.jl:27-44."""
    rng = np.random.default_rng(10)
    n, k = 50, 8
    Q, _ = np.linalg.qr(rng.standard_normal((n, k)))
    b = [Q[:, j].copy() for j in range(k)]
    x = rng.standard_normal(5)
    beta, v, nu = ko.householder_vec(x, 2)
    assert nu > 0 and abs(nu - np.linalg.norm(x)) < 1e-14
    Hm = np.eye(5) - beta * np.outer(v, v)
    y = Hm @ x
    assert abs(y[2] - nu) < 1e-13 and np.abs(np.delete(y, 2)).max() < 1e-13
    r = [1, 2, 3, 5, 6]
    ko.householder_rmul_basis(b, beta, v, r)
    Qd = Q.copy()
    Qd[:, r] = Q[:, r] @ Hm
    np.testing.assert_allclose(np.column_stack(b), Qd, atol=1e-13)
    c, s, rr = ko.givens(3.0, 4.0)
    assert abs(rr - 5.0) < 1e-14 and abs(c * 3 + s * 4 - 5) < 1e-14 and abs(-s * 3 + c * 4) < 1e-14


def test_stencil_spectrum_and_splitmix():
    A = ko.stencil_matrix(7, 5, 3, coeffs=(6.0, -1.0, -1.0, -1.0, -1.0, -1.0, -1.0))
    lam = ko.laplace_eigenvalues(7, 5, 3)
    np.testing.assert_allclose(np.linalg.eigvalsh(A.toarray()), lam, atol=1e-12)
    x = ko.splitmix_vector(1, 1000)
    assert 0 <= x.min() and x.max() < 1 and abs(x.mean() - 0.5) < 0.05
    assert np.array_equal(ko.splitmix_vector(1, 10, offset=5), ko.splitmix_vector(1, 15)[5:])


def test_cg_oracle():
    """test/linsolve.jl:1-117 (CG): b = A x for SPD A, shifted operator, numops = numiter + 1."""
    rng = np.random.default_rng(11)
    n = 100
    B = rng.standard_normal((n, n))
    A = B @ B.T / n + np.eye(n)
    b = rng.standard_normal(n)
    x, info = ko.linsolve_cg(A, b, maxiter=10 * n, tol=1e-12 * np.linalg.norm(b))
    assert info["converged"] == 1 and info["numops"] == info["numiter"] + 1
    np.testing.assert_allclose(A @ x, b, atol=1e-9)
    x, info = ko.linsolve_cg(A, b, maxiter=10 * n, tol=1e-10, a0=0.4, a1=1.7)
    assert info["converged"] == 1
    np.testing.assert_allclose(0.4 * x + 1.7 * (A @ x), b, atol=1e-8)
    x, info = ko.linsolve_cg(A, b, maxiter=3, tol=1e-14)
    assert info["converged"] == 0 and info["numiter"] == 3
    np.testing.assert_allclose(b - A @ x, info["residual"], atol=1e-10)


def test_bicgstab_oracle():
    """test/linsolve.jl:287-403 (BiCGStab), real scalar types: the small problem
    A = I - 0.9 B/ρ(B) (converged, b ≈ A x, restart from x costs one operator application,
    shifted operator), the large problem with maxiter = 2 (b ≈ (α₀ + α₁A) x + residual) and its
    continuation to convergence."""
    rng = np.random.default_rng(5)
    n, N = 10, 100
    B = rng.random((n, n)) - 0.5
    A = np.eye(n) - 0.9 * B / np.max(np.abs(np.linalg.eigvals(B)))
    b = rng.random(n)
    tol = 1e-12 * np.linalg.norm(b)
    x, info = ko.linsolve_bicgstab(A, b, maxiter=4 * n, tol=tol)
    assert info["converged"] > 0
    np.testing.assert_allclose(A @ x, b, rtol=1e-9)
    _, info2 = ko.linsolve_bicgstab(A, b, x, maxiter=4 * n, tol=tol)
    assert info2["numops"] == 1
    a0, a1 = rng.random() + 1, rng.random()
    x, info = ko.linsolve_bicgstab(A, b, maxiter=4 * n, tol=tol, a0=a0, a1=a1)
    assert info["converged"] > 0
    np.testing.assert_allclose(a0 * x + a1 * (A @ x), b, rtol=1e-9)

    A = rng.random((N, N)) - 0.5
    b = rng.random(N)
    a0, a1 = np.max(np.abs(np.linalg.eigvals(A))), -0.9 * rng.random()
    tol = 1e-12 * np.linalg.norm(b)
    x, info = ko.linsolve_bicgstab(A, b, maxiter=2, tol=tol, a0=a0, a1=a1)
    np.testing.assert_allclose(a0 * x + a1 * (A @ x) + info["residual"], b, rtol=1e-10)
    assert info["converged"] == 0 and info["numiter"] == 2
    x, info = ko.linsolve_bicgstab(A, b, x, maxiter=10 * N, tol=tol, a0=a0, a1=a1)
    assert info["converged"] > 0
    np.testing.assert_allclose(a0 * x + a1 * (A @ x), b, rtol=1e-9)
    # agrees with a direct solve
    np.testing.assert_allclose(x, np.linalg.solve(a0 * np.eye(N) + a1 * A, b), rtol=1e-8)
    # breakdown: r_shadow ⟂ r cannot happen at the first step (ρ = ‖r‖²) unless r = 0, which
    # returns before the loop with numops = 1
    x, info = ko.linsolve_bicgstab(np.eye(3), np.zeros(3), maxiter=5, tol=1e-12)
    assert info["converged"] == 1 and info["numops"] == 1 and info["numiter"] == 0


def test_lsmr_oracle():
    """test/lssolve.jl (LSMR), real scalar types: rank-deficient 2n×n problem — three iterations
    without reorthogonalisation leave residual = b − A x and normres = ‖Aᵀ r‖; with krylovdim = n it
    converges within n iterations to the minimum-norm solution V S⁺ Uᵀ b; with λ the regularised normal
    equations hold; the 2N×N problem with a ring of 5 vectors converges."""
    rng = np.random.default_rng(9)
    n, N = 10, 100
    A = rng.random((2 * n, n))
    U, S, Vt = np.linalg.svd(A, full_matrices=False)
    invS = 1 / S
    S[-1] = 0
    invS[-1] = 0
    A = U @ np.diag(S) @ Vt
    b = rng.random(2 * n)
    tol = 10 * n * np.finfo(float).eps
    x, info = ko.lssolve_lsmr(A, b, maxiter=3, krylovdim=1)
    r = b - A @ x
    np.testing.assert_allclose(info["residual"], r, rtol=1e-10)
    np.testing.assert_allclose(info["normres"], np.linalg.norm(A.T @ r), rtol=1e-8)
    assert info["converged"] == 0 and info["numops"] == 1 + 2 * 3
    x, info = ko.lssolve_lsmr(A, b, maxiter=n, tol=tol, krylovdim=n)
    assert info["converged"] > 0
    assert abs(Vt[-1] @ x) < tol
    np.testing.assert_allclose(x, Vt.T @ np.diag(invS) @ U.T @ b, rtol=1e-8)
    lam = rng.random()
    x, info = ko.lssolve_lsmr(A, b, maxiter=n, tol=tol, krylovdim=n, lam=lam)
    assert info["converged"] > 0
    np.testing.assert_allclose(A.T @ (b - A @ x), lam ** 2 * x, atol=2 * tol)
    A = rng.random((2 * N, N)) - 0.5
    b = rng.random(2 * N) - 0.5
    tol = 10 * N * np.finfo(float).eps
    x, info = ko.lssolve_lsmr(A, b, maxiter=N, tol=tol, krylovdim=5)
    assert info["converged"] > 0
    assert np.linalg.norm(A.T @ (b - A @ x)) < 5 * tol
    np.testing.assert_allclose(x, np.linalg.lstsq(A, b, rcond=None)[0], rtol=1e-8)


def _mat_with_eigrepetition(rng, N, mult):
    """test/testsetup.jl:46-58."""
    Q, _ = np.linalg.qr(rng.standard_normal((N, N)))
    D = np.sort(rng.standard_normal(N))
    i = 0
    while mult >= 2 and i + mult <= N // 2:
        D[i:i + mult] = D[i]
        D[N - i - mult:N - i] = D[N - i - 1]
        i += mult
        mult -= 1
    A = Q @ np.diag(D) @ Q.T
    return (A + A.T) / 2


def test_blocklanczos_oracle():
    """The reference's own BlockLanczos assertions (test/eigsolve.jl:552-770) on the restatement:
    both ends of a spectrum with repeated eigenvalues, orthonormal eigenvectors, block size 1 =
    Lanczos (same convergence count and restarts, one extra operator application, same residual
    norms), and a restart improving on a single sweep."""
    rng = np.random.default_rng(3)
    n, N, tol = 10, 100, 1e-12
    A = _mat_with_eigrepetition(rng, n, 2)
    X0 = [rng.random(n) for _ in range(2)]
    n1 = n // 2
    n2 = n - n1
    ev = np.linalg.eigvalsh(A)
    D1, V1, _ = ko.eigsolve_blocklanczos(A, X0, n1, "SR", krylovdim=n, maxiter=1, tol=tol)
    D2, V2, _ = ko.eigsolve_blocklanczos(A, X0, n2, "LR", krylovdim=2 * n, maxiter=4, tol=tol)
    np.testing.assert_allclose(np.concatenate([D1[:n1], D2[:n2][::-1]]), ev, rtol=1e-9, atol=1e-11)
    U1 = np.array(V1).T
    np.testing.assert_allclose(U1.T @ U1, np.eye(U1.shape[1]), atol=1e-10)
    np.testing.assert_allclose(A @ U1, U1 * D1, atol=1e-10)

    A = _mat_with_eigrepetition(rng, N, 4)
    X0 = [rng.random(N) for _ in range(4)]
    ev = np.linalg.eigvalsh(A)
    D1, _, i1 = ko.eigsolve_blocklanczos(A, X0, n, "SR", krylovdim=N, maxiter=10, tol=tol, eager=True)
    D2, _, i2 = ko.eigsolve_blocklanczos(A, X0, n, "LR", krylovdim=N, maxiter=10, tol=tol, eager=True)
    l1, l2 = i1["converged"], i2["converged"]
    assert l1 >= n and l2 >= n
    np.testing.assert_allclose(D1[:l1], ev[:l1], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(D2[:l2], ev[::-1][:l2], rtol=1e-9, atol=1e-10)

    A = rng.random((2 * N, 2 * N)) - 0.5
    A = (A + A.T) / 2
    x0 = rng.random(2 * N)
    e1, _, j1 = ko.eigsolve_lanczos(A, x0, n, "SR", krylovdim=2 * n, maxiter=10, tol=tol, orth=ko.Orth(ko.MGS2))
    e2, _, j2 = ko.eigsolve_blocklanczos(A, [x0], n, "SR", krylovdim=2 * n, maxiter=10, tol=tol)
    assert j1["converged"] == j2["converged"] and j1["numiter"] == j2["numiter"]
    assert j1["numops"] + 1 == j2["numops"]
    m = len(j1["normres"])
    np.testing.assert_allclose(j1["normres"], j2["normres"][:m], atol=1e-11)
    np.testing.assert_allclose(e1[:j1["converged"]], e2[:j2["converged"]], rtol=1e-10)

    A = _mat_with_eigrepetition(rng, N, 5)
    X0 = [rng.random(N) for _ in range(5)]
    v0 = np.linalg.eigvalsh(A)[:n]
    va, _, _ = ko.eigsolve_blocklanczos(A, X0, n, "SR", krylovdim=3 * n // 2, maxiter=1, tol=1e-12)
    vb, _, _ = ko.eigsolve_blocklanczos(A, X0, n, "SR", krylovdim=3 * n // 2, maxiter=2, tol=1e-12)
    assert np.linalg.norm(vb[:n // 2] - v0[:n // 2]) < np.linalg.norm(va[:n // 2] - v0[:n // 2])


def test_blocklanczos_toric_oracle():
    """test/eigsolve.jl:534-548: a block of 5 finds the four-fold degenerate −16 of the 3×3 toric code."""
    H = ko.toric_code_hamiltonian(3, 3)
    rng = np.random.default_rng(2)
    X0 = [rng.random(H.shape[0]) for _ in range(5)]
    D, _, info = ko.eigsolve_blocklanczos((-H).tocsr(), X0, 10, "SR", krylovdim=100, maxiter=1, tol=1e-6)
    assert np.sum(np.abs(D[:10] + 16.0) < 2.0 - 1e-6) == 4
    assert np.sum(np.abs(D[:10] + 16.0) < 1e-6) == 4


def _sort_real_then_imag(D):
    """sort(sort(D; by = imag, rev = true); alg = MergeSort, by = real) — test/eigsolve.jl:209."""
    D = np.asarray(D)
    D = D[np.argsort(-D.imag, kind="stable")]
    return D[np.argsort(D.real, kind="stable")]


@pytest.mark.parametrize("orth", [ko.Orth(ko.CGS2), ko.Orth(ko.MGS2), ko.Orth(ko.CGSIR, 0.75), ko.Orth(ko.MGSIR, 0.75)])
def test_arnoldi_eigsolve_oracle(orth):
    """test/eigsolve.jl:138-300 (Arnoldi, real scalar type): SR half + LR half reproduce eigvals(A),
    A V = V D for the complex eigenvectors, partial Schur relation for schursolve; iterative runs with
    restarts converge to the extremal eigenvalues and satisfy A V = V D + R."""
    rng = np.random.default_rng(7)
    n, N, tol = 10, 100, 1e-12
    A = rng.random((n, n)) - 0.5
    v = rng.random(n)
    n1 = n // 2
    n2 = n - n1
    D1, V1, i1 = ko.eigsolve_arnoldi(A, v, n1, "SR", krylovdim=n, maxiter=1, tol=tol, orth=orth)
    D2, V2, i2 = ko.eigsolve_arnoldi(A, v, n2, "LR", krylovdim=2 * n, maxiter=1, tol=tol, orth=orth)
    D = _sort_real_then_imag(np.linalg.eigvals(A))
    D2s = _sort_real_then_imag(D2)
    np.testing.assert_allclose(np.concatenate([D1[:n1], D2s[len(D2s) - n2:]]), D, rtol=1e-9, atol=1e-11)
    U1, U2 = np.column_stack(V1), np.column_stack(V2)
    np.testing.assert_allclose(A @ U1, U1 * D1, atol=1e-9)
    np.testing.assert_allclose(A @ U2, U2 * D2, atol=1e-9)
    T, Vs, vals, info = ko.schursolve_arnoldi(A, v, n1, "SR", krylovdim=n, maxiter=1, tol=tol, orth=orth)
    Q = np.column_stack(Vs)
    np.testing.assert_allclose(Q.T @ Q, np.eye(Q.shape[1]), atol=1e-10)
    np.testing.assert_allclose(A @ Q, Q @ T, atol=1e-9)
    with pytest.raises(ValueError):
        ko.eigsolve_arnoldi(A, v, n + 1, "LM", krylovdim=n, maxiter=1, tol=tol, orth=orth)

    A = rng.random((N, N)) - 0.5
    v = rng.random(N)
    Dfull = np.linalg.eigvals(A)
    Dfull = Dfull[np.argsort(-Dfull.imag, kind="stable")]
    for which, key in (("SR", lambda d: d.real), ("LR", lambda d: -d.real), ("LM", lambda d: -np.abs(d))):
        Dw, Vw, iw = ko.eigsolve_arnoldi(A, v, n, which, krylovdim=3 * n, maxiter=20, tol=tol, orth=orth, eager=True)
        l = iw["converged"]
        assert l > 0 and iw["numiter"] > 1
        want = Dfull[np.argsort(key(Dfull), kind="stable")][:l]
        if which == "LM":
            np.testing.assert_allclose(np.abs(Dw[:l]), np.abs(want), rtol=1e-8)
        else:
            np.testing.assert_allclose(Dw[:l], want, rtol=1e-8, atol=1e-10)
        Uw, Rw = np.column_stack(Vw), np.column_stack(iw["residual"])
        np.testing.assert_allclose(A @ Uw, Uw * Dw + Rw, atol=1e-9)


def _phi(A, v, p):
    """test/expintegrator.jl:1-13."""
    from scipy.linalg import expm
    m = A.shape[0]
    if p == 0:
        return expm(A) @ v
    Ap = np.zeros((m + p, m + p))
    Ap[:m, :m] = A
    Ap[:m, m] = v
    for k in range(1, p):
        Ap[m + k - 1, m + k] = 1
    return expm(Ap)[:m, -1]


@pytest.mark.parametrize("method", ["lanczos", "arnoldi"])
def test_expintegrator_oracle(method):
    """test/expintegrator.jl:15-190, real t: exponentiate = exp(A) column by column; expintegrator =
    Σ_j t^j ϕ_j(tA) u_j for p = 1..5, full-space and restarted; tol = 1e-3 stays within 1e-2|t|."""
    from scipy.linalg import expm
    rng = np.random.default_rng(13)
    n, N = 10, 100
    for orth in (ko.Orth(ko.CGS2), ko.Orth(ko.MGS2), ko.Orth(ko.CGSIR, 0.75), ko.Orth(ko.MGSIR, 0.75)):
        A = rng.random((n, n)) - 0.5
        if method == "lanczos":
            A = (A + A.T) / 2
        W = np.column_stack([ko.expintegrator(A, 1.0, (np.eye(n)[:, k],), method, orth, krylovdim=n, maxiter=2,
                                              tol=1e-12)[0] for k in range(n)])
        np.testing.assert_allclose(W, expm(A), rtol=1e-9, atol=1e-11)
        for t in (rng.random(), -rng.random()):
            for p in range(1, 6):
                u = tuple(rng.random(n) for _ in range(p + 1))
                w, info = ko.expintegrator(A, t, u, method, orth, krylovdim=n, maxiter=2, tol=1e-12)
                w2 = expm(t * A) @ u[0]
                for j in range(1, p + 1):
                    w2 = w2 + t ** j * _phi(t * A, u[j], j)
                assert info["converged"] > 0
                np.testing.assert_allclose(w, w2, rtol=1e-9, atol=1e-11)
    A = 0.5 * (rng.random((N, N)) - 0.5)
    if method == "lanczos":
        A = (A + A.T) / 2
    restarts = 0
    for t in (0.9 + rng.random(), -0.9 - rng.random()):
        for p in range(1, 6):
            u = tuple(rng.random(N) for _ in range(p + 1))
            w, info = ko.expintegrator(A, t, u, method, ko.Orth(ko.MGS2), krylovdim=n, maxiter=100, tol=1e-12, eager=True)
            w2 = expm(t * A) @ u[0]
            for j in range(1, p + 1):
                w2 = w2 + t ** j * _phi(t * A, u[j], j)
            assert info["converged"] > 0
            np.testing.assert_allclose(w, w2, rtol=1e-8, atol=1e-10)
            restarts += info["numiter"] - 1
            w, info = ko.expintegrator(A, t, u, method, ko.Orth(ko.MGS2), krylovdim=n, maxiter=100, tol=1e-3, eager=True)
            np.testing.assert_allclose(w, w2, atol=1e-2 * abs(t))
    assert restarts > 0


def test_issue133_lssolve_identity_fixture():
    """test/issues.jl:21-29 (issue #133), a literal known-answer fixture for LSMR."""
    x, info = ko.lssolve_lsmr(np.eye(2), np.array([1.0, 0.0]))
    assert np.array_equal(x, [1.0, 0.0])
    assert info["converged"] == 1 and info["numiter"] == 1 and info["numops"] == 2 and info["normres"] == 0.0


def test_realeigsolve_oracle():
    """test/eigsolve.jl:330-440: A = V D V⁻¹ with real spectrum — real eigenvalues in the requested order,
    real eigenvectors with A U = U D + R; a genuinely complex pair is flattened and reported."""
    from scipy.linalg import expm
    rng = np.random.default_rng(5)
    n, N = 10, 100
    V = expm(rng.standard_normal((N, N)) / 10)
    D = rng.standard_normal(N)
    A = V @ np.diag(D) @ np.linalg.inv(V)
    v = rng.random(N)
    for which, want in (("SR", np.sort(D)), ("LR", np.sort(D)[::-1]), ("LM", D[np.argsort(-np.abs(D))])):
        D1, V1, i1 = ko.realeigsolve_arnoldi(A, v, n, which, krylovdim=3 * n, maxiter=20, tol=1e-12, eager=True)
        l = i1["converged"]
        assert l > 0 and D1.dtype == np.float64 and not i1["ignored_imag"]
        np.testing.assert_allclose(D1[:l], want[:l], rtol=1e-8, atol=1e-10)
        U1, R1 = np.column_stack(V1), np.column_stack(i1["residual"])
        assert U1.dtype == np.float64
        np.testing.assert_allclose(A @ U1, U1 * D1 + R1, atol=1e-9)
    D1, _, i1 = ko.realeigsolve_arnoldi(np.array([[1.0, -1.0], [1.0, 1.0]]), np.array([1.0, 0.3]), 1, "LM", tol=1e-8)
    np.testing.assert_allclose(D1, 1.0)
    np.testing.assert_allclose(i1["ignored_imag"], [1.0])


def test_native_cpu_kernels_match_numpy_oracle():
    """oracle/csrc/kernels.c (the multi-threaded CPU baseline's n-length loops) against the numpy primitives
    they replace, then through the whole eigsolve driver: same numops, Ritz values to 1e-11."""
    from oracle import native
    if not native.available():
        pytest.skip("no gcc / OpenMP runtime")
    rng = np.random.default_rng(4)
    n, k = 211 * 237, 7
    x, y = rng.random(n), rng.random(n)
    B = [rng.random(n) for _ in range(k)]
    c = rng.standard_normal(k)
    h0, h1 = rng.random(k), None
    A = ko.stencil_matrix(211, 237)
    U = np.linalg.qr(rng.standard_normal((k, k)))[0][:, :4]
    want = dict(inner=ko.inner(x, y), norm=ko.norm(x), proj=ko.project(h0.copy(), B, x, 0.7, 0.3),
                unproj=ko.unproject(y, B, c, -1.0, 1.0), unproj0=ko.unproject(y, B, c, 2.0, 0.0),
                bt=ko.basistransform([b.copy() for b in B], U)[:4], mv=A @ x)
    with native.patched():
        assert ko.inner is native.inner
        np.testing.assert_allclose(ko.inner(x, y), want["inner"], rtol=1e-13)
        np.testing.assert_allclose(ko.norm(x), want["norm"], rtol=1e-13)
        np.testing.assert_allclose(ko.project(h0.copy(), B, x, 0.7, 0.3), want["proj"], rtol=1e-12)
        y_in = y.copy()
        np.testing.assert_allclose(ko.unproject(y, B, c, -1.0, 1.0), want["unproj"], rtol=1e-12, atol=1e-12)
        np.testing.assert_array_equal(y, y_in)                       # callers keep their input
        np.testing.assert_allclose(ko.unproject(y, B, c, 2.0, 0.0), want["unproj0"], rtol=1e-12, atol=1e-12)
        got = ko.basistransform([b.copy() for b in B], U)[:4]
        for g, w in zip(got, want["bt"]):
            np.testing.assert_allclose(g, w, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(native.CSR(A) @ x, want["mv"], rtol=1e-13)
        z = y.copy()
        ko._axpy(z, 0.25, x)
        np.testing.assert_allclose(z, y + 0.25 * x, rtol=1e-15)
        # float32 data falls through to numpy
        assert isinstance(ko.inner(x.astype(np.float32), y.astype(np.float32)), float)
    assert ko.inner is not native.inner
    nx, ny = 120, 90
    A = ko.stencil_matrix(nx, ny)
    x0 = ko.splitmix_vector(3, nx * ny)
    D0, _, i0 = ko.eigsolve_lanczos(A, x0, 3, "SR", krylovdim=16, maxiter=6, tol=0.0, orth=ko.Orth(ko.CGS2))
    with native.patched():
        D1, _, i1 = ko.eigsolve_lanczos(native.CSR(A), x0, 3, "SR", krylovdim=16, maxiter=6, tol=0.0, orth=ko.Orth(ko.CGS2))
    assert i0["numops"] == i1["numops"] and i0["converged"] == i1["converged"]
    np.testing.assert_allclose(D1, D0, rtol=1e-11)


def test_oracle_agrees_with_scipys_independent_implementations():
    """Pins from implementations that share no code with the oracle (or with KrylovKit): SciPy's LSMR
    (Fong & Saunders' own algorithm) iterate-for-iterate incl. the ‖Aᵀr‖ estimate, its BiCGStab
    iterate-for-iterate, expm_multiply (Al-Mohy & Higham) for exponentiate, ARPACK for the extremal
    eigenvalues and singular values."""
    import scipy.sparse.linalg as spl
    rng = np.random.default_rng(0)
    A = sp.random(300, 120, density=0.05, random_state=1).tocsr() + sp.eye(300, 120).tocsr()
    b = rng.random(300)
    for k in (3, 10, 25):
        x, info = ko.lssolve_lsmr(A.toarray(), b, maxiter=k, krylovdim=1, tol=0.0)
        ref = spl.lsmr(A, b, atol=0, btol=0, conlim=0, maxiter=k)
        np.testing.assert_allclose(x, ref[0], rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(info["normres"], ref[4], rtol=1e-9)            # ‖Aᵀ r‖ estimate
    x, _ = ko.lssolve_lsmr(A.toarray(), b, maxiter=15, krylovdim=1, tol=0.0, lam=0.3)
    np.testing.assert_allclose(x, spl.lsmr(A, b, damp=0.3, atol=0, btol=0, conlim=0, maxiter=15)[0], rtol=1e-5, atol=1e-8)

    N = sp.diags([-1.3, 2.6, -0.7], [-1, 0, 1], shape=(400, 400)).tocsr()
    bb = rng.random(400)
    x, _ = ko.linsolve_bicgstab(N, bb, maxiter=8, tol=1e-300)
    try:
        xs, _ = spl.bicgstab(N, bb, rtol=0, atol=0, maxiter=8)
    except TypeError:                                                              # older SciPy keyword
        xs, _ = spl.bicgstab(N, bb, tol=0, atol=0, maxiter=8)
    np.testing.assert_allclose(x, xs, rtol=1e-12, atol=1e-14)

    S = ko.stencil_matrix(30, 20).tocsr()
    v = rng.random(600)
    w, info = ko.expintegrator(S, -0.7, (v,), "lanczos", ko.Orth(ko.MGS2), krylovdim=30, tol=1e-12)
    assert info["converged"] == 1
    np.testing.assert_allclose(w, spl.expm_multiply(-0.7 * S, v), rtol=1e-11, atol=1e-13)
    C = (S + sp.diags([0.4], [1], shape=S.shape)).tocsr()                          # non-symmetric: Arnoldi
    w, info = ko.expintegrator(C, 0.35, (v,), "arnoldi", ko.Orth(ko.MGS2), krylovdim=30, tol=1e-12)
    np.testing.assert_allclose(w, spl.expm_multiply(0.35 * C, v), rtol=1e-10, atol=1e-12)

    vals, _, info = ko.eigsolve_lanczos(S, v, 3, "SR", krylovdim=30, maxiter=200, tol=1e-11, orth=ko.Orth(ko.MGS2))
    assert info["converged"] >= 3
    np.testing.assert_allclose(vals[:3], np.sort(spl.eigsh(S, k=3, which="SA", tol=1e-12)[0]), rtol=1e-9)
    D = rng.standard_normal((90, 40))
    sv, _, _, info = ko.svdsolve_gkl(D, rng.random(90), 3, "LR", krylovdim=20, maxiter=200, tol=1e-11, orth=ko.Orth(ko.MGS2))
    np.testing.assert_allclose(sv[:3], np.sort(spl.svds(D, k=3, tol=1e-12)[1])[::-1], rtol=1e-9)
