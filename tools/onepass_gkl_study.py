"""CPU study behind DESIGN.md §6 "not built: the one-pass-over-A GKL step" (SURVEY §8f-4, config 4).

The one-pass step streams A once per GKL step: while forming r' = A v_k it also accumulates z = A'(r'), and the next
step's A'u_{k+1} is recovered as (z - sum_j c_j A'u_j) / beta_k, where c are the coefficients removed from r'
(alpha_k for u_k plus the reorthogonalisation coefficients).  This script measures, in Float32 on a matrix of
config 4's shape and distribution, what that difference of O(sigma^2) quantities costs:

  * per step: || g_onepass - A'u || / || A'u ||  against the same quantity for the direct Float32 product,
    both measured against a Float64 product with the SAME u;
  * after `steps` steps (no restart): the Ritz values of the bidiagonal against a Float64 run of the standard step.

    python tools/onepass_gkl_study.py [m] [n] [steps]        (defaults 2000000 512 30; ~13 GB of RAM, a few minutes)
    python tools/onepass_gkl_study.py --restarted [m] [n]    the whole svdsolve of config 4 to convergence through the
                                                             oracle's driver, standard against one-pass operator

Test infrastructure / analysis only: numpy on the host, no part of the product path.
"""
import sys
import time

import numpy as np


def mgs2(r, basis, dtype):
    """two modified Gram-Schmidt sweeps of r against the columns of `basis`; returns the summed coefficients"""
    c = np.zeros(len(basis), dtype=np.float64)
    for _ in range(2):
        for j, q in enumerate(basis):
            s = dtype(np.dot(q, r))
            r -= s * q
            c[j] += float(s)
    return c


def gkl(A, u0, steps, dtype, onepass, A64=None):
    """GKL with full MGS2 reorthogonalisation of both sides (gkl.jl:324-346), standard or one-pass.
    Returns (alphas, betas, per-step relative errors of the A'u actually used, of the direct f32 product)."""
    A = A.astype(dtype, copy=False)
    U, V, G = [], [], []
    alphas, betas, err_used, err_direct = [], [], [], []
    u = (u0 / np.linalg.norm(u0)).astype(dtype)
    g = A.T @ u                                   # A'u_1: the first step has no predecessor to recycle
    v_prev, beta_prev = None, dtype(0)
    for k in range(steps):
        U.append(u)
        G.append(g)
        if A64 is not None:
            g64 = A64.T @ u.astype(np.float64)
            nrm = np.linalg.norm(g64)
            err_used.append(float(np.linalg.norm(g.astype(np.float64) - g64) / nrm))
            err_direct.append(float(np.linalg.norm((A.T @ u).astype(np.float64) - g64) / nrm))
        v = g.copy() if v_prev is None else g - beta_prev * v_prev
        mgs2(v, V, dtype)
        alpha = dtype(np.linalg.norm(v))
        v = v / alpha
        V.append(v)
        rp = A @ v                                # r' = A v_k   (the pass over A)
        z = A.T @ rp if onepass else None         # ... during which the one-pass kernel also forms A'(A v_k)
        r = rp - alpha * u
        c = mgs2(r, U, dtype)
        c[k] += float(alpha)                      # everything that was removed from r'
        beta = dtype(np.linalg.norm(r))
        u = r / beta
        if onepass:
            acc = z.astype(dtype).copy()
            for j in range(k + 1):
                acc -= dtype(c[j]) * G[j]
            g = acc / beta
        else:
            g = A.T @ u
        alphas.append(float(alpha))
        betas.append(float(beta))
        v_prev, beta_prev = v, beta
    return np.array(alphas), np.array(betas), np.array(err_used), np.array(err_direct)


def ritz(alphas, betas):
    k = len(alphas)
    B = np.diag(alphas)
    for i in range(k - 1):
        B[i + 1, i] = betas[i]
    return np.linalg.svd(B, compute_uv=False)


class OnePassOperator:
    """(apply_normal, apply_adjoint) pair for oracle/krylov_oracle.py's svdsolve driver that serves A'u WITHOUT a
    pass over A whenever u lies in the span of vectors whose images are known: every r' = A v comes with
    z = A'(r') from the same pass, every answered u is remembered with its image, and the restart's rotations of U
    are just other combinations of those.  The combination is found by least squares in Float64 (what the real
    step has exactly: alpha_k, the reorthogonalisation coefficients, the rotation matrices); the image is then
    accumulated in the working precision, like the kernel would."""

    def __init__(self, A, dtype, keep=96):
        self.A, self.dtype, self.keep = A, dtype, keep
        self.K, self.G = [], []                     # known vectors (length m) and their images A'k (length n)
        self.direct = self.recycled = self.normal = 0

    def _remember(self, k, g):
        self.K.append(k)
        self.G.append(g)
        if len(self.K) > self.keep:
            self.K.pop(0)
            self.G.pop(0)

    def apply_normal(self, v):
        self.normal += 1
        rp = self.A @ v
        self._remember(rp, (self.A.T @ rp).astype(self.dtype))     # the same pass over A in the one-pass kernel
        return rp.copy()

    def apply_adjoint(self, u):
        if self.K:
            Kt = np.array([k.astype(np.float64) for k in self.K])   # rows = known vectors
            u64 = u.astype(np.float64)
            gram = Kt @ Kt.T
            rhs = Kt @ u64
            c, *_ = np.linalg.lstsq(gram, rhs, rcond=1e-13)
            res = np.linalg.norm(u64 - Kt.T @ c) / np.linalg.norm(u64)
            if res < 3e-5:                                         # in the span (to Float32 working precision)
                g = np.zeros_like(self.G[0])
                for cj, gj in zip(c, self.G):
                    g += self.dtype(cj) * gj
                self.recycled += 1
                self._remember(u.copy(), g)
                return g.copy()
        self.direct += 1
        g = (self.A.T @ u).astype(self.dtype)
        self._remember(u.copy(), g)
        return g.copy()


def restarted(m, n):
    """svdsolve(GKL) to convergence (config 4's parameters: 6 triplets :LR, krylovdim 30, tol 1e-5) through the
    oracle's driver: standard operator against the one-pass operator, converged sigma against the Float64 truth."""
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import krylov_oracle as ko
    rng = np.random.default_rng(20260923)
    A32 = rng.random((m, n), dtype=np.float32) - np.float32(0.5)
    u0 = rng.random(m).astype(np.float32)
    t0 = time.time()
    G64 = np.zeros((n, n))
    for i in range(0, m, 100_000):                                 # A'A in Float64, blockwise
        blk = A32[i:i + 100_000].astype(np.float64)
        G64 += blk.T @ blk
    truth = np.sqrt(np.sort(np.linalg.eigvalsh(G64))[::-1][:6])
    print(f"# A {m} x {n} Float32; Float64 truth (sqrt eig A'A, {time.time() - t0:.0f} s): {np.array2string(truth, precision=6)}")
    for oname, orth in (("MGS2 (reference default)", ko.Orth(ko.MGS2)), ("CGSIR(eta = 0.75)", ko.Orth(ko.CGSIR, 0.75))):
        for name in ("standard", "one-pass"):
            t1 = time.time()
            if name == "standard":
                op = A32
            else:
                opo = OnePassOperator(A32, np.float32)
                op = (opo.apply_normal, opo.apply_adjoint)
            S, _, _, info = ko.svdsolve_gkl(op, u0.copy(), 6, "LR", krylovdim=30, maxiter=100, tol=1e-5, orth=orth)
            rel = np.abs(np.array(S[:6], dtype=np.float64) - truth) / truth
            extra = "" if name == "standard" else (f"; passes over A: {opo.normal} (each serving A v and A'(A v)) + {opo.direct} direct A'u, "
                                                   f"{opo.recycled} A'u recycled")
            print(f"{oname:26s} {name:9s}: converged {info['converged']}, numiter {info['numiter']}, numops {info['numops']}, "
                  f"sigma_1..6 max rel err vs truth {rel.max():.2e}, normres max {np.max(np.abs(info['normres'][:6])):.1e}"
                  f"{extra}  [{time.time() - t1:.0f} s]")


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--restarted":
        return restarted(int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000, int(sys.argv[3]) if len(sys.argv) > 3 else 512)
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    rng = np.random.default_rng(20260923)
    t0 = time.time()
    A32 = rng.random((m, n), dtype=np.float32) - np.float32(0.5)      # config 4's distribution: uniform(-0.5, 0.5)
    A64 = A32.astype(np.float64)
    u0 = rng.random(m)
    print(f"# A {m} x {n} Float32 uniform(-0.5, 0.5), {steps} GKL steps with MGS2 on both sides; generated in {time.time() - t0:.0f} s")
    a64, b64, _, _ = gkl(A64, u0, steps, np.float64, False)
    s64 = ritz(a64, b64)
    print(f"# Float64 standard step: alpha ~ {a64.mean():.1f}, beta ~ {b64.mean():.2f} (beta/alpha ~ {b64.mean() / a64.mean():.0f}); "
          f"Ritz values {s64[0]:.6f} .. {s64[5]:.6f}")
    for name, op in (("standard (two passes)", False), ("one-pass", True)):
        t1 = time.time()
        a, b, eu, ed = gkl(A32, u0, steps, np.float32, op, A64)
        s = ritz(a, b)
        rel = np.abs(s[:6] - s64[:6]) / s64[:6]
        print(f"{name:22s} Float32: ||A'u used - A'u(f64)|| / ||A'u||  median {np.median(eu[1:]):.2e}  max {eu[1:].max():.2e}   "
              f"(direct f32 product: median {np.median(ed[1:]):.2e});  Ritz values 1..6 vs the Float64 run: max rel {rel.max():.2e}  "
              f"[{time.time() - t1:.0f} s]")
        print(f"{'':22s} sigma_1..6 = {np.array2string(s[:6], precision=5)}")


if __name__ == "__main__":
    main()
