"""world_size-2 CPU (gloo) test of the N>1 path's host logic: the row sharding, the halo
plan and the column localisation of krylovkit.jl_b200/sharding.py drive a sharded
Lanczos-CGS2 recurrence (local SpMV with exchanged halos + all-reduced coefficients, the
communication pattern of SURVEY §8e) that must reproduce the serial oracle."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, nx, ny, steps, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from krylovkit_jl_b200 import sharding as sh
    from oracle import krylov_oracle as ko

    A = ko.stencil_matrix(nx, ny)
    n = nx * ny
    shard = sh.shard_grid_lines(nx, ny, rank, world)
    r0, r1 = shard.row_offset, shard.row_offset + shard.n_local
    Aloc = A[r0:r1].tocsr()
    cmin, cmax = int(Aloc.indices.min()), int(Aloc.indices.max())
    nl = torch.zeros(world, dtype=torch.int64)
    nl[rank] = shard.n_local
    dist.all_reduce(nl)
    lo, hi, ok = sh.halo_plan(shard, cmin, cmax, nl.tolist())
    assert ok and lo == (nx if rank > 0 else 0) and hi == (nx if rank < world - 1 else 0)
    loc = np.array([sh.localize_column(int(g), shard, lo) for g in Aloc.indices], dtype=np.int64)
    import scipy.sparse as sp
    Aext = sp.csr_matrix((Aloc.data, loc, Aloc.indptr), shape=(shard.n_local, shard.n_local + lo + hi))
    # neighbours' needs = my sends
    need = torch.zeros(2 * world, dtype=torch.int64)
    need[2 * rank], need[2 * rank + 1] = lo, hi
    dist.all_reduce(need)
    send_up = int(need[2 * (rank + 1)]) if rank + 1 < world else 0      # rank+1's lo halo
    send_dn = int(need[2 * (rank - 1) + 1]) if rank > 0 else 0          # rank-1's hi halo

    def matvec(x):
        halo_lo, halo_hi = torch.zeros(lo, dtype=torch.float64), torch.zeros(hi, dtype=torch.float64)
        reqs = []
        if send_up:
            reqs.append(dist.isend(torch.from_numpy(x[-send_up:].copy()), rank + 1))
        if send_dn:
            reqs.append(dist.isend(torch.from_numpy(x[:send_dn].copy()), rank - 1))
        if lo:
            reqs.append(dist.irecv(halo_lo, rank - 1))
        if hi:
            reqs.append(dist.irecv(halo_hi, rank + 1))
        for q in reqs:
            q.wait()
        return Aext @ np.concatenate([x, halo_lo.numpy(), halo_hi.numpy()])

    def allsum(v):
        t = torch.tensor(np.atleast_1d(np.asarray(v, dtype=np.float64)))
        dist.all_reduce(t)
        return t.numpy()

    # sharded Lanczos with ClassicalGramSchmidt2 (lanczos.jl:180-222, 313-324)
    x0 = ko.splitmix_vector(20260923, n)[r0:r1]
    b0 = np.sqrt(allsum(x0 @ x0)[0])
    Ax = matvec(x0)
    alpha = allsum(x0 @ Ax)[0] / (b0 * b0)
    v = x0 / b0
    r = Ax / b0 - alpha * v
    d = allsum(v @ r)[0]
    alpha += d
    r = r - d * v
    beta = np.sqrt(allsum(r @ r)[0])
    V, alphas, betas = [v], [alpha], [beta]
    for _ in range(steps):
        V.append(r / betas[-1])
        w = matvec(V[-1])
        a = allsum(V[-1] @ w)[0]                       # AllReduce(1)
        w = w - betas[-1] * V[-2]
        w = w - a * V[-1]
        s = allsum([q @ w for q in V])                 # AllReduce(k): the h-coefficients
        for q, sj in zip(V, s):
            w = w - sj * q
        a += s[-1]
        alphas.append(a)
        betas.append(np.sqrt(allsum(w @ w)[0]))        # AllReduce(1)
        r = w
    if rank == 0:
        np.save(out, np.array([alphas, betas]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_lanczos_matches_serial_oracle(tmp_path):
    import torch.multiprocessing as mp
    from oracle import krylov_oracle as ko
    nx, ny, steps, world = 40, 33, 12, 2
    out = str(tmp_path / "ab.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, nx, ny, steps, out), nprocs=world, join=True)
    ab = np.load(out)
    A = ko.stencil_matrix(nx, ny)
    f = ko.lanczos_initialize(A, ko.splitmix_vector(20260923, nx * ny), ko.Orth(ko.CGS2))
    for _ in range(steps):
        f = ko.lanczos_expand(A, f, ko.Orth(ko.CGS2))
    np.testing.assert_allclose(ab[0], f.alphas, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(ab[1], f.betas, rtol=1e-12, atol=1e-13)


def test_shards_partition_rows():
    from krylovkit_jl_b200 import sharding as sh
    for world in (1, 2, 3, 4, 8):
        shards = [sh.shard_grid_lines(4000, 2500, r, world) for r in range(world)]
        assert shards[0].row_offset == 0
        assert sum(s.n_local for s in shards) == 4000 * 2500
        for a, b in zip(shards, shards[1:]):
            assert a.row_offset + a.n_local == b.row_offset
        assert all(s.n_local % 4000 == 0 for s in shards)
    with pytest.raises(ValueError):
        sh.shard_grid_lines(10, 10, 2, 2)
