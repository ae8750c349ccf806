"""Thread-level model of k_dense_onepass (index arithmetic only): every thread's loads, smem stores, shuffles."""
import numpy as np
OP_T, OP_ROWS, OP_PAD, OP_UNR = 256, 32, 33, 8
def run(m, n, dtype, grid):
    T = np.dtype(dtype)
    VEC = 16 // T.itemsize; VPC = OP_ROWS // VEC; CSTEP = OP_T // VPC
    NZ = 1 if n <= 256 else 2 if n <= 512 else 4 if n <= 1024 else 7
    ld = (m + 31) // 32 * 32
    rng = np.random.default_rng(1)
    A = np.zeros((n, ld), dtype=np.float64); A[:, :m] = rng.random((n, m)) - 0.5     # A[c, i] column-major storage
    x = rng.random(n) - 0.5
    y = np.full(m, np.nan); ntiles = ld // OP_ROWS
    zpart = np.zeros((grid, n))
    for b in range(grid):
        zacc = np.zeros((OP_T, NZ))
        for tile in range(b, ntiles, grid):
            i0 = tile * OP_ROWS
            As = np.full(n * OP_PAD, np.nan); written = np.zeros(n * OP_PAD, int)
            acc = np.zeros((OP_T, VEC))
            for tid in range(OP_T):
                rq, c0 = tid % VPC, tid // VPC
                cb = c0
                while cb < n:
                    for u in range(OP_UNR):
                        c = cb + u * CSTEP
                        if c < n:
                            a = A[c, i0 + rq * VEC: i0 + rq * VEC + VEC]
                            for e in range(VEC):
                                acc[tid, e] += a[e] * x[c]
                                As[c * OP_PAD + rq * VEC + e] = a[e]; written[c * OP_PAD + rq * VEC + e] += 1
                    cb += CSTEP * OP_UNR
            # every (c,row) written exactly once
            for c in range(n):
                assert (written[c*OP_PAD:c*OP_PAD+32] == 1).all() and written[c*OP_PAD+32] == 0
            # shuffles
            off = VPC
            while off < 32:
                new = acc.copy()
                for tid in range(OP_T):
                    lane, warp = tid & 31, tid >> 5
                    new[tid] = acc[tid] + acc[warp * 32 + (lane ^ off)]
                acc = new; off <<= 1
            ys_part = np.full((8, 32), np.nan)
            for tid in range(OP_T):
                lane, warp = tid & 31, tid >> 5
                if lane < VPC:
                    for e in range(VEC): ys_part[warp, lane * VEC + e] = acc[tid, e]
            ys = ys_part.sum(axis=0)
            for t in range(32):
                if i0 + t < m: y[i0 + t] = ys[t]
            for tid in range(OP_T):
                for s in range(NZ):
                    c = s * OP_T + tid
                    if c < n:
                        zacc[tid, s] += sum(As[c * OP_PAD + row] * ys[row] for row in range(32))
        for tid in range(OP_T):
            for s in range(NZ):
                c = s * OP_T + tid
                if c < n: zpart[b, c] = zacc[tid, s]
    # reduce kernel model
    z = np.zeros(n)
    for c in range(n):
        red = np.zeros(8)
        for g in range(8):
            s = [0.0] * 4; p = g
            while p + 24 < grid:
                for q in range(4): s[q] += zpart[p + 8 * q, c]
                p += 32
            while p < grid:
                s[0] += zpart[p, c]; p += 8
            red[g] = (s[0] + s[1]) + (s[2] + s[3])
        z[c] = red.sum()
    Am = A[:, :m].T
    yr = Am @ x
    assert np.allclose(y, yr, rtol=1e-12, atol=1e-13), (m, n, dtype)
    assert np.allclose(z, Am.T @ yr, rtol=1e-11, atol=1e-12), (m, n, dtype)
    print("ok", m, n, dtype, grid)
for (m, n, g) in [(100, 70, 3), (64, 512, 2), (33, 300, 5), (1000, 17, 40), (96, 1030, 1), (4000, 6, 37)]:
    for dt in ("f4", "f8"):
        run(m, n, dt, g)
