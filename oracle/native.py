"""TEST / BASELINE INFRASTRUCTURE ONLY — plugs the OpenMP kernels of oracle/csrc/kernels.c into the
oracle's primitives so that the CPU baseline runs multi-threaded like KrylovKit's own `Array` fast path
(src/orthonormal.jl:66-73, 151-196, 322-354) instead of at numpy / scipy single-thread speed.

    with native.patched():                       # swaps ko.apply / inner / norm / _axpy / project /
        ko.eigsolve_lanczos(A, x0, ...)          # unproject / basistransform for float64 data

The drivers (restart logic, convergence tests, Householder bookkeeping) remain the numpy oracle's; only the
n-length loops change.  Only tests/, bench.py's CPU legs and __graft_entry__ import this module.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os

import numpy as np
import scipy.sparse as sp

from . import krylov_oracle as ko
from .csrc import build as _build

_lib = None
_P = C.POINTER


def load():
    global _lib
    if _lib is None:
        os.environ.setdefault("OMP_WAIT_POLICY", "passive")     # only effective if libgomp is not loaded yet
        lib = C.CDLL(_build.build())
        dp, ip32, ip64 = _P(C.c_double), _P(C.c_int32), _P(C.c_int64)
        pp = _P(dp)
        lib.nat_num_threads.restype = C.c_int
        lib.nat_set_num_threads.argtypes = [C.c_int]
        lib.nat_set_num_threads.restype = None
        lib.nat_spmv_csr_f64.argtypes = [C.c_int64, ip64, ip32, dp, dp, dp]
        lib.nat_dot_f64.restype = C.c_double
        lib.nat_dot_f64.argtypes = [C.c_int64, dp, dp]
        lib.nat_axpy_f64.argtypes = [C.c_int64, C.c_double, dp, dp]
        lib.nat_scale_f64.argtypes = [C.c_int64, C.c_double, dp, dp]
        lib.nat_project_f64.argtypes = [C.c_int64, C.c_int32, pp, dp, C.c_double, C.c_double, dp]
        lib.nat_unproject_f64.argtypes = [C.c_int64, C.c_int32, pp, dp, C.c_double, C.c_double, dp]
        lib.nat_basistransform_f64.argtypes = [C.c_int64, C.c_int32, C.c_int32, pp, dp, pp]
        for f in ("nat_spmv_csr_f64", "nat_axpy_f64", "nat_scale_f64", "nat_project_f64", "nat_unproject_f64",
                  "nat_basistransform_f64"):
            getattr(lib, f).restype = None
        _lib = lib
    return _lib


def num_threads() -> int:
    return int(load().nat_num_threads())


def set_num_threads(t: int) -> None:
    load().nat_set_num_threads(int(t))


def calibrate(A: "CSR", x: np.ndarray, k: int = 8, verbose=None) -> int:
    """Pick the OpenMP thread count that is actually fastest on this host for one representative step
    (fresh-output matvec + project + in-place unproject + dot at the real n): on shared / quota-limited
    boxes `nproc` threads can be far slower than fewer.  Leaves the winner installed and returns it."""
    import time
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cands = sorted({t for t in (ncpu, ncpu // 2, ncpu // 4, ncpu // 8, 16, 8, 4) if 1 <= t <= ncpu}, reverse=True)
    basis = [x * (1.0 + 0.01 * j) for j in range(k)]
    h = np.zeros(k)
    best_t, best = cands[0], float("inf")
    for t in cands:
        set_num_threads(t)
        dt = float("inf")
        for _ in range(2):
            t0 = time.perf_counter()
            w = A @ x
            project(h, basis, w)
            w = unproject(w, basis, h, -1.0, 1.0, inplace=True)
            inner(w, x)
            dt = min(dt, time.perf_counter() - t0)
        if verbose:
            verbose(f"native.calibrate: {t} threads -> {dt * 1e3:.1f} ms per step sample")
        if dt < best:
            best_t, best = t, dt
    set_num_threads(best_t)
    return best_t


def _d(a):
    return a.ctypes.data_as(_P(C.c_double))


def _ok(*arrs) -> bool:
    return all(isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous and a.ndim == 1
               for a in arrs)


def _ptrs(vecs):
    arr = (_P(C.c_double) * len(vecs))()
    for i, v in enumerate(vecs):
        arr[i] = _d(v)
    return arr


class CSR:
    """CSR operator with the index arrays converted once (int64 rowptr, int32 colidx)."""

    def __init__(self, A):
        A = sp.csr_matrix(A)
        A.sort_indices()
        self.shape = A.shape
        self.rowptr = np.ascontiguousarray(A.indptr, dtype=np.int64)
        self.colidx = np.ascontiguousarray(A.indices, dtype=np.int32)
        self.vals = np.ascontiguousarray(A.data, dtype=np.float64)
        self.scipy = A

    def __matmul__(self, x):
        if not _ok(x):
            return self.scipy @ x
        y = np.empty(self.shape[0])
        load().nat_spmv_csr_f64(self.shape[0], self.rowptr.ctypes.data_as(_P(C.c_int64)),
                                self.colidx.ctypes.data_as(_P(C.c_int32)), _d(self.vals), _d(x), _d(y))
        return y


def inner(x, y):
    if _ok(x, y):
        return float(load().nat_dot_f64(len(x), _d(x), _d(y)))
    return _orig["inner"](x, y)


def norm(x):
    if _ok(x):
        return float(np.sqrt(load().nat_dot_f64(len(x), _d(x), _d(x))))
    return _orig["norm"](x)


def _axpy(y, a, x):
    if _ok(x, y):
        load().nat_axpy_f64(len(y), float(a), _d(x), _d(y))
        return y
    return _orig["_axpy"](y, a, x)


def project(y, b, x, alpha=1.0, beta=0.0, r=None):
    vecs = list(b) if r is None else [b[i] for i in r]
    if vecs and _ok(x, *vecs) and isinstance(y, np.ndarray) and y.dtype == np.float64 and y.flags.c_contiguous:
        load().nat_project_f64(len(x), len(vecs), _ptrs(vecs), _d(x), float(alpha), float(beta), _d(y))
        return y
    return _orig["project"](y, b, x, alpha, beta, r)


def scale_(x, a):
    if _ok(x):
        load().nat_scale_f64(len(x), float(a), _d(x), _d(x))
        return x
    return _orig["scale_"](x, a)


def unproject(y, b, x, alpha=1.0, beta=0.0, r=None, inplace=False):
    vecs = list(b) if r is None else [b[i] for i in r]
    if vecs and _ok(y, *vecs):
        if beta == 0:
            out = np.empty_like(y)
        elif inplace and beta == 1:
            out = y                                             # the caller owns y (unproject!! semantics)
        else:
            out = y.copy()                                      # the oracle's callers keep their input
        c = np.ascontiguousarray(x, dtype=np.float64)
        load().nat_unproject_f64(len(out), len(vecs), _ptrs(vecs), _d(c), float(alpha), float(beta), _d(out))
        return out
    return _orig["unproject"](y, b, x, alpha, beta, r, inplace)


def basistransform(b, U):
    m, keep = U.shape
    if m == len(b) and _ok(*b[:m]):
        Uc = np.asfortranarray(U, dtype=np.float64)
        out = [np.empty_like(b[0]) for _ in range(keep)]
        load().nat_basistransform_f64(len(b[0]), m, keep, _ptrs(b[:m]), Uc.ctypes.data_as(_P(C.c_double)), _ptrs(out))
        for j in range(keep):
            b[j] = out[j]
        return b
    return _orig["basistransform"](b, U)


_NAMES = ("inner", "norm", "_axpy", "scale_", "project", "unproject", "basistransform")
_orig = {name: getattr(ko, name) for name in _NAMES}


@contextlib.contextmanager
def patched():
    """Route the oracle's n-length primitives through the OpenMP kernels (float64 data; anything else
    falls through to numpy).  Wrap sparse operators in `native.CSR` to get the threaded matvec too."""
    load()
    for name in _NAMES:
        setattr(ko, name, globals()[name])
    try:
        yield
    finally:
        for name in _NAMES:
            setattr(ko, name, _orig[name])


def available() -> bool:
    try:
        load()
        return True
    except (OSError, FileNotFoundError, Exception):      # no gcc / no OpenMP runtime: numpy oracle only
        return False
