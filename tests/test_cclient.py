"""The drop-in boundary is a C-ABI: a plain C99 program (tests/cclient/lanczos_client.c) includes
include/b200krylov.h, links libb200krylov.so and drives Lanczos steps with no Python in between — what a
Julia `ccall` / cgo / JNI binding does.  CPU tier: it compiles with -std=c99 -pedantic, links every symbol
it uses, and fails loudly without a device.  GPU tier: its coefficients equal the oracle's."""
import os
import subprocess

import numpy as np
import pytest

from conftest import HAVE_GPU, ROOT
from oracle import krylov_oracle as ko

SRC = os.path.join(ROOT, "tests", "cclient", "lanczos_client.c")
LIBDIR = os.path.join(ROOT, "krylovkit.jl_b200")


SRC2 = os.path.join(ROOT, "tests", "cclient", "restart_client.c")


def _build(tmp_path, src=SRC):
    exe = str(tmp_path / os.path.basename(src)[:-2])
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
           "-o", exe, src, "-L", LIBDIR, "-lb200krylov", "-lm", f"-Wl,-rpath,{LIBDIR}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_header_is_plain_c_and_client_links(tmp_path):
    _build(tmp_path, SRC2)
    exe = _build(tmp_path)
    if HAVE_GPU:
        return
    r = subprocess.run([exe, "10", "10", "3", "2"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 2                       # loud failure, not a silent CPU path
    assert "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("tag,otag", [(2, ko.CGS2), (3, ko.MGS2), (0, ko.CGS)])
def test_c_client_lanczos_matches_oracle(tmp_path, tag, otag):
    exe = _build(tmp_path)
    nx, ny, steps = 301, 211, 24
    r = subprocess.run([exe, str(nx), str(ny), str(steps), str(tag)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    ab = np.array([[float(t) for t in ln.split()] for ln in lines[:steps]])
    A = ko.stencil_matrix(nx, ny)
    x0 = ko.splitmix_vector(20260923, nx * ny)
    orth = ko.Orth(otag)
    f = ko.lanczos_initialize(A, x0, orth)
    for _ in range(steps - 1):
        f = ko.lanczos_expand(A, f, orth)
    tol = 1e-11 if otag != ko.CGS else 1e-6       # plain CGS loses orthogonality; coefficients drift with it
    np.testing.assert_allclose(ab[:, 0], f.alphas, rtol=tol)
    np.testing.assert_allclose(ab[:, 1], f.betas, rtol=tol)
    defect = float(lines[steps].split()[1])
    launches = int(lines[steps].split()[3])
    assert launches > steps
    if otag != ko.CGS:
        assert defect < 1e-12


@pytest.mark.gpu
def test_c_client_thick_restart_cycle_matches_oracle(tmp_path):
    """A C program drives b2k_lanczos_expand_many (device-chained steps) -> b2k_host_lanczos_restart ->
    b2k_basis_transform -> shrink -> b2k_lanczos_expand_many through one whole thick restart: the Ritz values
    equal the oracle's eigsolve(maxiter = 2, tol = 0) on the same (A, x0).  (The client diagonalises T with its
    own Jacobi sweep, so eigenvector signs differ from LAPACK's: same Krylov space, same Ritz values.)"""
    exe = _build(tmp_path, SRC2)
    nx, ny, kd, hm = 301, 211, 30, 3
    r = subprocess.run([exe, str(nx), str(ny), str(kd), str(hm)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    A = ko.stencil_matrix(nx, ny)
    x0 = ko.splitmix_vector(20260923, nx * ny)
    vals, _, info = ko.eigsolve_lanczos(A, x0, hm, "SR", krylovdim=kd, maxiter=2, tol=0.0, orth=ko.Orth(ko.CGS2))
    assert int(lines[0].split()[1]) == info["numops"]
    got = np.array([[float(t) for t in ln.split()] for ln in lines[1:1 + hm]])
    np.testing.assert_allclose(got[:, 0], vals[:hm], rtol=1e-9)
    np.testing.assert_allclose(got[:, 1], info["normres"][:hm], rtol=1e-5)
    assert int(lines[1 + hm].split()[1]) > 0
