"""exponentiate / expintegrator — mirror of src/matrixfun/{exponentiate,expintegrator}.jl.

y(t) = ϕ₀(tA) u₀ + t ϕ₁(tA) u₁ + … + tᵖ ϕₚ(tA) uₚ by adaptive Krylov time stepping: the device does the
Lanczos / Arnoldi expansions, one tall-skinny combination per accepted step and a handful of axpys; the
(K+p+1)² matrix exponential and the step-size control stay on the host.  Real `t` only (no complex
dtype in the library): imaginary-time evolution yes, real-time Schrödinger evolution no.
"""
from __future__ import annotations

import math
import warnings

import numpy as np
from scipy.linalg import expm

from .algorithms import Arnoldi, ConvergenceInfo, Lanczos, WARN_LEVEL
from .factorizations import arnoldi as ar
from .factorizations import lanczos as lz
from .operators import apply
from .orthonormal import unproject_
from .vectors import B200Vec


def exponentiate(A, t: float, v: B200Vec, alg: Lanczos | Arnoldi | None = None, **kwargs):
    """exponentiate(A, t, v, alg) = expintegrator(A, t, (v,), alg) — exponentiate.jl:83-84."""
    return expintegrator(A, t, (v,), alg, **kwargs)


def expintegrator(A, t: float, u, alg: Lanczos | Arnoldi | None = None, **kwargs):
    """expintegrator(A, t, u::Tuple, alg::Union{Lanczos,Arnoldi}) — expintegrator.jl:101-323.
    `alg.tol` is the requested accuracy per unit time.  Without `alg`: Lanczos if `issymmetric=True`
    is passed, else Arnoldi (expintegrator.jl:88-99)."""
    if isinstance(u, B200Vec):
        u = (u,)
    u = tuple(u)
    if alg is None:
        sym = kwargs.pop("issymmetric", False) or kwargs.pop("ishermitian", False)
        alg = Lanczos(**kwargs) if sym else Arnoldi(**kwargs)
    if isinstance(t, complex):
        if t.imag != 0:
            raise TypeError("expintegrator: complex time steps need complex vectors, which libb200krylov does not have")
        t = t.real
    t = float(t)
    if len(u) == 1:
        u = (u[0], u[0].zerovector())
    p = len(u) - 1
    lanczos = isinstance(alg, Lanczos)
    fz = lz if lanczos else ar
    u0 = u[0]
    Au0 = apply(A, u0)
    numops = 1
    w0 = u0.copy()
    maxiter, krylovdim = alg.maxiter, alg.krylovdim
    assert maxiter >= 1
    eta = float(alg.tol)                          # tolerance per unit time
    totalerr = 0.0
    sgn = float(np.sign(t))
    tau = abs(t)
    if math.isfinite(tau):
        dtau, dtaumin, maxerr = tau, tau / maxiter, tau * eta
    else:
        dtau, dtaumin, maxerr = 1.0, 0.0, eta
    gamma = 0.8                                   # safety factor (δ = 1.2 is unused in the reference too)
    tau0 = 0.0
    w: list = [None] * (p + 1)
    w[0] = w0
    w[1] = Au0                                    # the reference copies Au₀; nothing else holds it here

    def refill(reuse_first: bool):
        """w[j] = A w[j-1] + Σ_l (sgn τ₀)^l / l! · u[j+l]  — expintegrator.jl:145-156, 283-292."""
        nonlocal numops
        for j in range(1, p + 1):
            if j > 1 or not reuse_first:
                w[j] = apply(A, w[j - 1])
                numops += 1
            lfac = 1
            for l in range(0, p - j + 1):
                w[j] = w[j].add_(u[j + l], (sgn * tau0) ** l / lfac)
                lfac *= l + 1

    refill(True)
    beta = w[p].norm()
    if beta < eta and p == 1:                     # u₀ is a fixed point of the ODE
        return w0, ConvergenceInfo(1, None, beta, 0, numops)
    Iter = lz.LanczosIterator if lanczos else ar.ArnoldiIterator
    it = Iter(A, w[p], alg.orth)
    fact = fz.initialize(it)
    numops += 1
    numiter = 1

    def rayleigh(K):
        if lanczos:
            dv, ev = fact.rayleighquotient()
            return np.diag(dv[:K]) + np.diag(ev[:K - 1], 1) + np.diag(ev[:K - 1], -1)
        return fact.rayleighquotient()

    def small_exp(K, dt):
        H = np.zeros((K + p + 1, K + p + 1))
        H[:K, :K] = rayleigh(K) * (sgn * dt)
        H[0, K] = 1
        for i in range(1, p + 1):
            H[K + i - 1, K + i] = 1
        return expm(H)

    def take_step(K, dt, expH):
        nonlocal w0
        jfac = 1
        for j in range(1, p):
            w0 = w0.add_(w[j], (sgn * dt) ** j / jfac)
            jfac *= j + 1
        w[p] = unproject_(w[p], fact.basis(), np.ascontiguousarray(expH[:K, K + p - 1]))
        w[p] = w[p].add_(fact.residual(), float(expH[K - 1, K + p]))        # first correction
        w0 = w0.add_(w[p], beta * (sgn * dt) ** p)
        w[0] = w0

    while True:
        K = len(fact)
        if K == krylovdim:
            if numiter < maxiter:
                dtau = min(dtau, tau - tau0)
                if math.isfinite(tau):            # adapt the minimal time step
                    dtaumin = (tau - tau0) / (maxiter - numiter + 1)
            else:
                dtau = tau - tau0
            expH = small_exp(K, dtau)
            eps_ = abs(dtau ** p * beta * fact.normres() * expH[K - 1, K + p])
            omega = eps_ / (dtau * eta)
            q = K / 2
            while numiter < maxiter and omega >= 1 and dtau > dtaumin:
                eps_prev, dtau_prev = eps_, dtau
                dtau = max(dtau * (gamma / omega) ** (1 / (q + 1)), dtaumin)
                expH = small_exp(K, dtau)
                eps_ = abs(dtau ** p * beta * fact.normres() * expH[K - 1, K + p])
                omega = eps_ / (dtau * eta)
                q = max(0.0, math.log(eps_ / eps_prev) / math.log(dtau / dtau_prev) - 1)
            tau0 = tau0 + dtau if numiter < maxiter else tau
            totalerr += eps_
            take_step(K, dtau, expH)
            if omega < gamma:                     # be more ambitious next time
                dtau *= (gamma / omega) ** (1 / (q + 1))
        elif fact.normres() <= (tau - tau0) * eta or alg.eager:
            dt = tau - tau0
            expH = small_exp(K, dt)
            eps_ = abs(dt ** p * beta * fact.normres() * expH[K - 1, K + p])
            omega = eps_ / (dt * eta)
            if omega < 1:
                totalerr += eps_
                take_step(K, dt, expH)
                tau0 = tau
        if tau0 >= tau:
            ok = totalerr <= maxerr
            if not ok and alg.verbosity >= WARN_LEVEL:
                warnings.warn(f"expintegrate did not reach sufficiently small error after {numiter} iterations: "
                              f"total error = {totalerr}, numops = {numops}")
            return w0, ConvergenceInfo(int(ok), None, totalerr, numiter, numops)
        if K < krylovdim:
            fact = fz.expand_(it, fact)
            numops += 1
        else:
            refill(False)
            beta = w[p].norm()
            if beta < eta and p == 1:             # w₀ is a fixed point of the ODE
                return w0, ConvergenceInfo(1, None, beta, numiter, numops)
            it = Iter(A, w[p], alg.orth)
            fact = fz.initialize_(it, fact)
            numops += 1
            numiter += 1
