"""Thread-level model of k_dense_onepass_w (variant B): 64-row tiles, 512 threads, register refill."""
import numpy as np
T_, ROWS, PAD, LD_ = 512, 64, 65, 16
def run(m, n, grid):
    VPC = ROWS // 4; CSTEP = T_ // VPC
    ld = (m + 31) // 32 * 32
    rng = np.random.default_rng(2)
    Aflat = np.zeros(n * ld + 64)      # a little slack to detect out-of-range reads via NaN
    Aflat[:] = np.nan
    A2 = np.zeros((n, ld)); A2[:, :m] = rng.random((n, m)) - 0.5
    Aflat[:n * ld] = A2.reshape(-1)
    x = rng.random(n) - 0.5
    ntiles = (ld + ROWS - 1) // ROWS
    grid = min(grid, ntiles)
    y = np.full(m, np.nan); zpart = np.zeros((grid, n))
    def ldv(tile, tid, u):
        rq, c0 = tid % VPC, tid // VPC
        c = c0 + u * CSTEP
        r0 = tile * ROWS + rq * 4
        ok = tile < ntiles and r0 < ld and c < n
        if not ok: return np.zeros(4)
        v = Aflat[c * ld + r0: c * ld + r0 + 4]
        assert not np.isnan(v).any()
        return v.copy()
    for b in range(grid):
        a = np.zeros((T_, LD_, 4))
        for tid in range(T_):
            for u in range(LD_): a[tid, u] = ldv(b, tid, u)
        zacc = np.zeros(T_)
        tile = b
        while tile < ntiles:
            i0 = tile * ROWS; nxt = tile + grid
            As = np.full(n * PAD, np.nan); acc = np.zeros((T_, 4))
            for tid in range(T_):
                rq, c0 = tid % VPC, tid // VPC
                for u in range(LD_):
                    c = c0 + u * CSTEP
                    if c < n:
                        for e in range(4):
                            acc[tid, e] += a[tid, u, e] * x[c]
                            As[c * PAD + rq * 4 + e] = a[tid, u, e]
                        a[tid, u] = ldv(nxt, tid, u)
            new = acc.copy()
            for tid in range(T_):
                lane, warp = tid & 31, tid >> 5
                new[tid] = acc[tid] + acc[warp * 32 + (lane ^ 16)]
            acc = new
            ys_part = np.full((16, 64), np.nan)
            for tid in range(T_):
                lane, warp = tid & 31, tid >> 5
                if lane < VPC:
                    for e in range(4): ys_part[warp, lane * 4 + e] = acc[tid, e]
            ys = ys_part.sum(axis=0)
            for t in range(64):
                if i0 + t < m: y[i0 + t] = ys[t]
            for tid in range(T_):
                if tid < n:
                    col = As[tid * PAD: tid * PAD + 64]
                    assert not np.isnan(col).any()
                    zacc[tid] += float(col @ ys)
            tile = nxt
        for tid in range(T_):
            if tid < n: zpart[b, tid] = zacc[tid]
    z = zpart.sum(axis=0)
    Am = A2[:, :m].T
    yr = Am @ x
    assert np.allclose(y, yr, rtol=1e-12, atol=1e-13)
    assert np.allclose(z, Am.T @ yr, rtol=1e-11, atol=1e-12)
    print("ok", m, n, grid)
for (m, n, g) in [(100, 70, 3), (64, 512, 2), (33, 300, 5), (1000, 17, 4), (96, 500, 1), (4000, 6, 7), (160, 512, 2)]:
    run(m, n, g)
