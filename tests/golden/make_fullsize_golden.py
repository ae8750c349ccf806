"""Generate tests/golden/fullsize.json: the ORACLE's results (oracle/krylov_oracle.py, the CPU restatement of
KrylovKit.jl; n-length loops in the OpenMP kernels of oracle/csrc/kernels.c) for the BASELINE.json configurations
at FULL size.  The GPU arm of bench.py and the `-m gpu` full-size parity tests compare against this file: the
GPU box has the oracle too, but a whole config-2 job costs ~20 s of its 16-CPU quota (config 3/5 more), so the
numbers are computed once here and committed.  Nothing in the product reads this file.

    python tests/golden/make_fullsize_golden.py [c2 c2_mgs2 c2_1e6 c3 c3_mgs2 c5s c5 c4_truth]     (default: all)

Inputs are the synthetic ones SURVEY.md §8d fixes (stencils + splitmix64 start vectors, seed 20260923), so the
file is reproducible bit for bit up to the summation order of the OpenMP reductions (<= 1e-13 relative).
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
os.environ.setdefault("OMP_WAIT_POLICY", "passive")

from oracle import krylov_oracle as ko          # noqa: E402
from oracle import native                       # noqa: E402

SEED = 20260923
OUT = os.path.join(HERE, "fullsize.json")
CONV = (4.0, -1.4, -0.6, -1.2, -0.8, 0, 0)       # convection-diffusion of config 3


def lanczos_case(nx, ny, nz, kd, cycles, orth, howmany=4):
    n = nx * ny * nz
    # 5-point (diag 4) in 2-D, 7-point (diag 6) in 3-D — SURVEY §8d configs 2 and 5
    A = native.CSR(ko.stencil_matrix(nx, ny, nz, (6.0, -1, -1, -1, -1, -1, -1)) if nz > 1 else ko.stencil_matrix(nx, ny))
    x0 = ko.splitmix_vector(SEED, n)
    out = {"grid": [nx, ny, nz], "n": n, "krylovdim": kd, "orth": orth, "howmany": howmany, "seed": SEED,
           "which": "SR", "tol": 0.0, "after_cycles": {}}
    o = ko.Orth({"cgs2": ko.CGS2, "mgs2": ko.MGS2}[orth])
    for c in cycles:
        t0 = time.perf_counter()
        with native.patched():
            vals, _, info = ko.eigsolve_lanczos(A, x0, howmany, "SR", krylovdim=kd, maxiter=c, tol=0.0, orth=o)
        out["after_cycles"][str(c)] = {"ritz": [float(v) for v in vals[:howmany]],
                                       "normres": [float(v) for v in info["normres"][:howmany]],
                                       "numops": int(info["numops"]), "numiter": int(info["numiter"]),
                                       "oracle_seconds": round(time.perf_counter() - t0, 1)}
        print(nx, ny, nz, orth, c, out["after_cycles"][str(c)], flush=True)
    return out


def gmres_case(nx, ny, kd, cycles, orth):
    n = nx * ny
    A = native.CSR(ko.stencil_matrix(nx, ny, 1, CONV))
    b = A @ np.ones(n)
    o = ko.Orth({"cgs2": ko.CGS2, "mgs2": ko.MGS2}[orth])
    out = {"grid": [nx, ny, 1], "n": n, "krylovdim": kd, "orth": orth, "coeffs": list(CONV), "b": "A*ones",
           "tol": 1e-300, "after_cycles": {}}
    idx = [0, 1, n // 3, n // 2, n - 2, n - 1]
    for c in cycles:
        t0 = time.perf_counter()
        with native.patched():
            x, info = ko.linsolve_gmres(A, b, None, krylovdim=kd, maxiter=c, tol=1e-300, orth=o)
        out["after_cycles"][str(c)] = {"normres": float(info["normres"]), "numops": int(info["numops"]),
                                       "numiter": int(info["numiter"]), "x_norm": float(np.linalg.norm(x)),
                                       "err_norm": float(np.linalg.norm(x - 1.0)),
                                       "x_samples": {str(i): float(x[i]) for i in idx},
                                       "oracle_seconds": round(time.perf_counter() - t0, 1)}
        print(nx, ny, orth, c, out["after_cycles"][str(c)], flush=True)
    return out


def dense_truth_case(m, n, howmany=6):
    """configs[3] (svdsolve on the dense 2e6 x 512 Float32 splitmix matrix): the largest singular values of the matrix
    the device generates, from the Float64 Gram matrix A'A accumulated blockwise (sqrt of its eigenvalues) — the
    truth the Float32 runs of bench.py's c4 / c4_onepass records are compared with (3e-5 relative)."""
    t0 = time.perf_counter()
    A = ko.dense_splitmix(SEED, m, n)
    G = np.zeros((n, n))
    for i in range(0, m, 100_000):
        blk = A[i:i + 100_000].astype(np.float64)
        G += blk.T @ blk
    sig = np.sqrt(np.sort(np.linalg.eigvalsh(G))[::-1][:howmany])
    out = {"shape": [m, n], "dtype": "float32", "seed": SEED, "howmany": howmany, "which": "LR",
           "sigma_float64_truth": [float(v) for v in sig], "oracle_seconds": round(time.perf_counter() - t0, 1)}
    print(m, n, out, flush=True)
    return out


CASES = {
    # BASELINE.json configs[1]: the bench workload (5 cycles) and the 2-cycle prefix the GPU test checks
    "c2": lambda: lanczos_case(4000, 2500, 1, 60, (2, 5), "cgs2"),
    "c2_mgs2": lambda: lanczos_case(4000, 2500, 1, 60, (2,), "mgs2"),          # the reference's default orthogonalizer
    "c2_1e6": lambda: lanczos_case(1000, 1000, 1, 60, (2, 5), "cgs2"),
    # configs[2]: GMRES(40) on the 1e7 convection-diffusion problem, 5 restart cycles
    "c3": lambda: gmres_case(4000, 2500, 40, (2, 5), "cgs2"),
    "c3_mgs2": lambda: gmres_case(4000, 2500, 40, (2,), "mgs2"),
    # configs[4] shape (7-point 3-D Laplacian, krylovdim 30) at a size one GPU test finishes quickly: n = 8e6
    "c5s": lambda: lanczos_case(200, 200, 200, 30, (2, 5), "cgs2"),
    # configs[4] itself: 8e7 rows (625 x 500 x 256), krylovdim 30, 3 restart cycles (the bench's other_configs record)
    "c5": lambda: lanczos_case(625, 500, 256, 30, (3,), "cgs2"),
    # configs[3]: the Float64 truth of the dense Float32 matrix's top singular values
    "c4_truth": lambda: dense_truth_case(2_000_000, 512),
}


def main():
    want = sys.argv[1:] or list(CASES)
    data = json.load(open(OUT)) if os.path.exists(OUT) else {}
    native.load()
    native.set_num_threads(len(os.sched_getaffinity(0)))
    for name in want:
        data[name] = CASES[name]()
        data["_meta"] = {"generator": "tests/golden/make_fullsize_golden.py", "oracle": "oracle/krylov_oracle.py + oracle/native.py",
                         "threads": native.num_threads(), "note": "oracle results, not reference-run results: no Julia in the image"}
        with open(OUT, "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
