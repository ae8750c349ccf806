"""Device context and device vectors: the VectorInterface side of the drop-in boundary.

`B200Vec` implements exactly the method set KrylovKit needs from a vector type — the set
`InnerProductVec` implements in the reference (src/innerproductvec.jl:82-137):
scalartype, zerovector, scale / scale!! (`scale_`), add!! (`add_`), inner, norm.  Julia's
`f!!` ("mutate if possible, always use the return value") is spelled `f_` here; every `_`
method mutates in place and returns the mutated object, which is what KrylovKit relies on
(lanczos.jl:257, orthonormal.jl:523).

All arithmetic happens in libb200krylov.so on the GPU; numpy only appears at the
upload/download boundary.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _lib as L


def _np_dtype(dtype_code: int):
    return np.float64 if dtype_code == L.F64 else np.float32


class B200Context:
    """One device context = one GPU, one stream, one or more vector spaces (slabs).

    ncols is the number of vector slots of space 0; a Krylov solver needs about
    krylovdim + 6 of them (the basis, the residual and a few work vectors).
    """

    def __init__(self, n_local: int, ncols: int, dtype=np.float64, device: int = 0,
                 rank: int = 0, nranks: int = 1, nccl_uid: bytes | None = None,
                 n_global: int | None = None, row_offset: int = 0):
        self.lib = L.load()
        self.np_dtype = np.dtype(dtype)
        if self.np_dtype == np.float64:
            self.dtype = L.F64
        elif self.np_dtype == np.float32:
            self.dtype = L.F32
        else:
            raise ValueError("B200Context: dtype must be float64 or float32 (real only)")
        self.n_local = int(n_local)
        self.rank, self.nranks = int(rank), int(nranks)
        self.n_global = int(n_global) if n_global is not None else int(n_local)
        self.row_offset = int(row_offset)
        h = L.c_ctx()
        if nranks > 1:
            if nccl_uid is None or len(nccl_uid) != 128:
                raise ValueError("B200Context: nccl_uid must be the 128-byte ncclUniqueId")
            buf = C.create_string_buffer(nccl_uid, 128)
            st = self.lib.b2k_ctx_create_dist(C.byref(h), device, self.n_local, ncols, self.dtype,
                                              rank, nranks, buf, self.n_global, self.row_offset)
        else:
            st = self.lib.b2k_ctx_create(C.byref(h), device, self.n_local, ncols, self.dtype)
        L.check(st, None)
        self.h = h
        self.space_n = [self.n_local]
        self._alive = [True]
        self._finalizer = weakref.finalize(self, _destroy_ctx, self.lib, h, self._alive)

    # -- plumbing -----------------------------------------------------------------
    def check(self, status: int):
        L.check(status, self.h)

    def close(self):
        self._finalizer()

    def sync(self):
        self.check(self.lib.b2k_ctx_sync(self.h))

    @property
    def launches(self) -> int:
        return int(self.lib.b2k_ctx_launch_count(self.h))

    @property
    def stream(self) -> int:
        return int(self.lib.b2k_ctx_stream(self.h) or 0)

    def add_space(self, n_local: int, ncols: int, sharded: bool = True) -> int:
        sp = C.c_int32()
        self.check(self.lib.b2k_space_create(self.h, n_local, ncols, 1 if sharded else 0, C.byref(sp)))
        self.space_n.append(int(n_local))
        return sp.value

    # -- vector construction ----------------------------------------------------
    def empty(self, space: int = 0) -> "B200Vec":
        v = L.c_vec()
        self.check(self.lib.b2k_vec_alloc(self.h, space, C.byref(v)))
        return B200Vec(self, v.value)

    def empty_range(self, count: int, space: int = 0) -> list["B200Vec"]:
        v = L.c_vec()
        self.check(self.lib.b2k_vec_alloc_range(self.h, space, count, C.byref(v)))
        return [B200Vec(self, v.value + i) for i in range(count)]

    def zeros(self, space: int = 0) -> "B200Vec":
        v = self.empty(space)
        self.check(self.lib.b2k_vec_zero(self.h, v.handle))
        return v

    def full(self, value: float, space: int = 0) -> "B200Vec":
        v = self.empty(space)
        self.check(self.lib.b2k_vec_fill(self.h, v.handle, float(value)))
        return v

    def from_host(self, x, space: int = 0) -> "B200Vec":
        v = self.empty(space)
        v.upload(x)
        return v

    def splitmix(self, seed: int, space: int = 0) -> "B200Vec":
        v = self.empty(space)
        self.check(self.lib.b2k_vec_fill_splitmix(self.h, v.handle, seed))
        return v


class B200Vec:
    """A vector resident in one column of a device slab (handle = space << 20 | column)."""

    __slots__ = ("ctx", "handle", "_fin", "__weakref__")

    def __init__(self, ctx: B200Context, handle: int):
        self.ctx = ctx
        self.handle = int(handle)
        # return the column to the slab when the Python object dies (Julia: finalizer)
        self._fin = weakref.finalize(self, _free_handle, ctx.lib, ctx.h, self.handle, ctx._alive)

    # -- bookkeeping ---------------------------------------------------------------
    @property
    def space(self) -> int:
        return self.handle >> 20

    def __len__(self) -> int:
        return self.ctx.space_n[self.space]

    def free(self):
        self._fin()

    def disown(self):
        """Forget the column without releasing it: the library has taken it back (b2k_lanczos_expand_many)."""
        self._fin.detach()

    def upload(self, x):
        a = np.ascontiguousarray(x, dtype=self.ctx.np_dtype)
        if a.shape != (len(self),):
            raise L.DimensionMismatch(f"upload: expected shape ({len(self)},), got {a.shape}")
        self.ctx.check(self.ctx.lib.b2k_vec_upload(self.ctx.h, self.handle, a.ctypes.data))
        return self

    def to_host(self, out: np.ndarray | None = None) -> np.ndarray:
        """Download into `out` (e.g. a pinned buffer) or a fresh array."""
        if out is None:
            out = np.empty(len(self), dtype=self.ctx.np_dtype)
        elif out.shape != (len(self),) or out.dtype != self.ctx.np_dtype or not out.flags.c_contiguous:
            raise L.DimensionMismatch("to_host: out has the wrong shape / dtype / layout")
        self.ctx.check(self.ctx.lib.b2k_vec_download(self.ctx.h, self.handle, out.ctypes.data))
        return out

    # -- VectorInterface -----------------------------------------------------------
    def scalartype(self):
        return self.ctx.np_dtype.type

    def zerovector(self) -> "B200Vec":
        """zerovector(v): a new zero vector in the same space."""
        return self.ctx.zeros(self.space)

    def zerovector_(self) -> "B200Vec":
        self.ctx.check(self.ctx.lib.b2k_vec_zero(self.ctx.h, self.handle))
        return self

    def copy(self) -> "B200Vec":
        w = self.ctx.empty(self.space)
        self.ctx.check(self.ctx.lib.b2k_vec_copy(self.ctx.h, w.handle, self.handle))
        return w

    def scale(self, alpha: float) -> "B200Vec":
        """scale(v, α): out of place."""
        w = self.ctx.empty(self.space)
        self.ctx.check(self.ctx.lib.b2k_vec_scale(self.ctx.h, w.handle, self.handle, float(alpha)))
        return w

    def scale_(self, alpha: float, src: "B200Vec | None" = None) -> "B200Vec":
        """scale!!(v, α) or scale!!(v, src, α): v <- α·src."""
        s = self if src is None else src
        self.ctx.check(self.ctx.lib.b2k_vec_scale(self.ctx.h, self.handle, s.handle, float(alpha)))
        return self

    def add_(self, w: "B200Vec", alpha: float = 1.0, beta: float = 1.0) -> "B200Vec":
        """add!!(v, w, α, β): v <- β·v + α·w."""
        self.ctx.check(self.ctx.lib.b2k_vec_axpby(self.ctx.h, self.handle, w.handle, float(alpha),
                                                  float(beta)))
        return self

    def add(self, w: "B200Vec", alpha: float = 1.0, beta: float = 1.0) -> "B200Vec":
        return self.copy().add_(w, alpha, beta)

    def inner(self, w: "B200Vec") -> float:
        """inner(v, w) (real: symmetric)."""
        out = C.c_double()
        self.ctx.check(self.ctx.lib.b2k_vec_inner(self.ctx.h, self.handle, w.handle, C.byref(out)))
        return out.value

    def norm(self) -> float:
        out = C.c_double()
        self.ctx.check(self.ctx.lib.b2k_vec_norm(self.ctx.h, self.handle, C.byref(out)))
        return out.value


def _free_handle(lib, ctx_h, handle, alive):
    if not alive[0]:
        return
    try:
        st = lib.b2k_vec_free(ctx_h, handle)
    except Exception:  # interpreter shutdown
        return
    if st != 0:        # a finalizer cannot raise usefully, but a failing free is a bookkeeping bug: say so
        import sys
        sys.stderr.write(f"b200krylov: b2k_vec_free(handle {handle:#x}) failed with status {st}\n")


def _destroy_ctx(lib, ctx_h, alive):
    if alive[0]:
        alive[0] = False
        lib.b2k_ctx_destroy(ctx_h)


def cache_release() -> None:
    """Give the library's cached device / pinned blocks back to the driver (b2k_cache_release).  Freed slabs
    and CSR arrays (>= 32 MB) are kept for reuse by the next context of the same shape, up to B2K_CACHE_GB
    (default 16 GB); call this when other CUDA libraries in the process need the memory."""
    L.check(L.load().b2k_cache_release())


# free functions in VectorInterface style -------------------------------------------------
def inner(v: B200Vec, w: B200Vec) -> float:
    return v.inner(w)


def norm(v: B200Vec) -> float:
    return v.norm()


def handles(vecs) -> "C.Array":
    arr = (L.c_vec * len(vecs))()
    for i, v in enumerate(vecs):
        arr[i] = v.handle
    return arr
