"""Kernel micro-benchmarks on one GPU (CUDA events on the context stream, inputs >> L2).

    python tools/microbench.py [--n 10000000] [--k 48] [--reps 20]

Prints one JSON object per kernel with achieved algorithmic GB/s and the fraction of the
measured HBM peak (MEASURED_PEAKS.json).  Not the bench contract — see bench.py.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import krylovkit_jl_b200 as kk  # noqa: E402
from krylovkit_jl_b200 import _lib as L  # noqa: E402

cudart = None


def _cudart():
    global cudart
    if cudart is None:
        for name in ("libcudart.so.12", "libcudart.so"):
            try:
                cudart = C.CDLL(name)
                break
            except OSError:
                continue
        if cudart is None:
            import glob
            cands = glob.glob("/usr/local/cuda/lib64/libcudart.so*")
            cudart = C.CDLL(cands[0])
    return cudart


class Timer:
    def __init__(self, stream):
        self.rt = _cudart()
        self.stream = C.c_void_p(stream)
        self.e0, self.e1 = C.c_void_p(), C.c_void_p()
        assert self.rt.cudaEventCreate(C.byref(self.e0)) == 0
        assert self.rt.cudaEventCreate(C.byref(self.e1)) == 0

    def time(self, fn, reps, warm=3):
        for _ in range(warm):
            fn()
        self.rt.cudaStreamSynchronize(self.stream)
        self.rt.cudaEventRecord(self.e0, self.stream)
        for _ in range(reps):
            fn()
        self.rt.cudaEventRecord(self.e1, self.stream)
        self.rt.cudaEventSynchronize(self.e1)
        ms = C.c_float()
        self.rt.cudaEventElapsedTime(C.byref(ms), self.e0, self.e1)
        return ms.value / reps


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--nx", type=int, default=4000)
    ap.add_argument("--k", type=int, default=48)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--coop", type=int, default=1)
    a = ap.parse_args()
    n, k = a.n, a.k
    ny = n // a.nx
    assert a.nx * ny == n
    pk, pk_kind = peak()
    ctx = kk.B200Context(n, k + 8)
    lib = ctx.lib
    lib.b2k_debug_set_coop(a.coop)
    tm = Timer(ctx.stream)
    W = 8 * n
    out = []

    def rec(name, ms, nbytes):
        gbs = nbytes / ms / 1e6
        out.append({"kernel": name, "ms": round(ms, 4), "GBs": round(gbs, 1),
                    "frac": round(gbs / pk, 3), "peak": pk, "peak_kind": pk_kind, "bytes": nbytes})
        print(json.dumps(out[-1]), flush=True)

    basis = ctx.empty_range(k)
    for j, v in enumerate(basis):
        lib.b2k_vec_fill_splitmix(ctx.h, v.handle, 1000 + j)
    b = kk.OrthonormalBasis(basis)
    x = ctx.splitmix(1)
    y = ctx.splitmix(2)
    hs = kk.vectors.handles(basis)
    d = C.c_double()
    h = (C.c_double * (k + 1))()

    # bring the clocks up first: a B200 idles at 120 MHz and needs tens of milliseconds of load to reach its
    # running clocks — r01's BLAS-1 figures (0.38-0.57) were taken cold and were clock ramp, not kernel quality
    # (tools/probes/blas1_probe.cu: the same launch shape runs at 0.83-0.87 once warm)
    nrm0, ps0 = C.c_double(), C.c_int32()
    for _ in range(30):
        lib.b2k_basis_orthogonalize(ctx.h, y.handle, hs, k, h, L.CGS2, 0.0, C.byref(nrm0), C.byref(ps0))
    lib.b2k_vec_fill_splitmix(ctx.h, y.handle, 2)
    # scalars-returning calls block on a D2H copy + stream sync per call (the KrylovKit contract): their time
    # includes that round trip; the *_async rows time the same kernels without it
    rec("vec_inner", tm.time(lambda: lib.b2k_vec_inner(ctx.h, x.handle, y.handle, C.byref(d)), a.reps), 2 * W)
    rec("vec_norm", tm.time(lambda: lib.b2k_vec_norm(ctx.h, x.handle, C.byref(d)), a.reps), W)
    rec("vec_axpby", tm.time(lambda: lib.b2k_vec_axpby(ctx.h, y.handle, x.handle, 1e-9, 1.0), a.reps), 3 * W)
    rec("vec_scale", tm.time(lambda: lib.b2k_vec_scale(ctx.h, y.handle, x.handle, 0.5), a.reps), 2 * W)

    op = kk.B200CSR.stencil(ctx, a.nx, ny)
    spmv_bytes = op.nnz * 12 + 4 * (n + 1) + 2 * W
    rec("spmv_csr", tm.time(lambda: lib.b2k_op_apply(ctx.h, op.h, x.handle, y.handle), a.reps), spmv_bytes)
    rec("spmv_csr_dot", tm.time(lambda: lib.b2k_op_apply_dot(ctx.h, op.h, x.handle, y.handle, x.handle,
                                                              C.byref(d)), a.reps), spmv_bytes + W)

    rec(f"project_k{k}", tm.time(lambda: lib.b2k_basis_project(ctx.h, hs, k, x.handle, 1.0, 0.0, h), a.reps),
        (k + 1) * W)
    c = (C.c_double * k)(*([1e-12] * k))
    rec(f"unproject_k{k}", tm.time(lambda: lib.b2k_basis_unproject(ctx.h, y.handle, hs, k, c, 1.0, 1.0),
                                   a.reps), (k + 2) * W)
    nrm = C.c_double()
    ps = C.c_int32()
    for alg, name, nb in ((L.CGS, "cgs", (2 * k + 3) * W), (L.CGS2, "cgs2_fused", (3 * k + 5) * W)):
        rec(f"{name}_k{k}", tm.time(lambda: lib.b2k_basis_orthogonalize(ctx.h, y.handle, hs, k, h, alg, 0.0,
                                                                     C.byref(nrm), C.byref(ps)), a.reps), nb)
    rec(f"mgs_k{k}", tm.time(lambda: lib.b2k_basis_orthogonalize(ctx.h, y.handle, hs, k, h, L.MGS, 0.0,
                                                                C.byref(nrm), C.byref(ps)), max(2, a.reps // 4)),
        (4 * k + 1) * W)
    keep = (3 * k) // 5
    U = np.linalg.qr(np.random.default_rng(0).standard_normal((k, k)))[0][:, :keep]
    Uf = np.asfortranarray(U)
    rec(f"basistransform_{k}x{keep}", tm.time(lambda: lib.b2k_basis_transform(
        ctx.h, hs, k, Uf.ctypes.data_as(C.POINTER(C.c_double)), k, keep), max(2, a.reps // 4)), (k + keep) * W)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"microbench_n{n}_k{k}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
