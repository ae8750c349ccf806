"""world_size-2 CPU (gloo) run of EVERY driver on a row-sharded context: the real host code of
krylovkit.jl_b200 (eigsolve Lanczos / Arnoldi / BlockLanczos, GMRES, CG, BiCGStab, exponentiate) runs on
two ranks against tests/hostsim.py in its sharded mode — local rows, all-reduced scalars (gloo standing in
for NCCL), gathered x standing in for the halo exchange.  What this proves is the property the multi-GPU
design rests on: the host logic is rank-replicated — every rank sees the same scalars, takes
the same branches, issues the same collectives in the same order — and the sharded results equal the
serial oracle's.  The same checks run on real GPUs in tools/dist_check.py (tests/test_gpu_zz_dist.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out, fused):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "2")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import hostsim
    import krylovkit_jl_b200 as kk
    from krylovkit_jl_b200 import sharding
    from oracle import krylov_oracle as ko

    nx, ny = 30, 22
    n = nx * ny
    shard = sharding.shard_grid_lines(nx, ny, rank, world)
    sl = slice(shard.row_offset, shard.row_offset + shard.n_local)

    def gather(vec):
        parts = [None] * world
        dist.all_gather_object(parts, vec.to_host())
        return np.concatenate(parts)

    report = {}
    with hostsim.installed(fused=fused):
        ctx = kk.B200Context(shard.n_local, 160, rank=rank, nranks=world, nccl_uid=bytes(128), n_global=n,
                             row_offset=shard.row_offset)
        A = ko.stencil_matrix(nx, ny)
        cdc = (5.0, -1.4, -0.6, -1.2, -0.8, 0.0, 0.0)
        Acd = ko.stencil_matrix(nx, ny, 1, cdc)
        op = kk.B200CSR.stencil(ctx, nx, ny)
        opcd = kk.B200CSR.stencil(ctx, nx, ny, 1, cdc)
        x0 = ko.splitmix_vector(11, n)
        assert np.array_equal(ctx.splitmix(11).to_host(), x0[sl])          # global counter RNG
        y = kk.apply(op, ctx.from_host(x0[sl]))
        np.testing.assert_allclose(gather(y), A @ x0, rtol=1e-14)
        # uploaded CSR with global column indices
        Aloc = Acd[sl].tocsr()
        op2 = kk.B200CSR.from_csr_arrays(ctx, shard.n_local, n, Aloc.indptr, Aloc.indices, Aloc.data)
        np.testing.assert_allclose(gather(kk.apply(op2, ctx.from_host(x0[sl]))), Acd @ x0, rtol=1e-14)

        # Lanczos eigsolve with restarts
        alg = kk.Lanczos(orth=kk.cgs2, krylovdim=20, maxiter=60, tol=1e-10, verbosity=0)
        vals, vecs, info = kk.eigsolve(op, ctx.from_host(x0[sl]), 2, "SR", alg)
        ovals, _, oinfo = ko.eigsolve_lanczos(A, x0, 2, "SR", krylovdim=20, maxiter=60, tol=1e-10, orth=ko.Orth(ko.CGS2))
        assert info.numops == oinfo["numops"] and info.numiter == oinfo["numiter"]
        np.testing.assert_allclose(vals[:2], ovals[:2], rtol=1e-10)
        v = gather(vecs[0])
        assert np.linalg.norm(A @ v - vals[0] * v) < 1e-8
        report["lanczos"] = [float(t) for t in vals[:2]]
        del vecs, info

        # GMRES, CG, BiCGStab
        b = ko.splitmix_vector(5, n)
        x, info = kk.linsolve(opcd, ctx.from_host(b[sl]), None, kk.GMRES(orth=kk.mgs2, krylovdim=15, maxiter=40, tol=1e-10, verbosity=0))
        ox, oinfo = ko.linsolve_gmres(Acd, b, krylovdim=15, maxiter=40, tol=1e-10, orth=ko.Orth(ko.MGS2))
        assert info.converged == 1 and info.numops == oinfo["numops"]
        np.testing.assert_allclose(gather(x), ox, rtol=1e-8, atol=1e-10)
        x, info = kk.linsolve(opcd, ctx.from_host(b[sl]), None, kk.BiCGStab(maxiter=200, tol=1e-10, verbosity=0))
        ox, oinfo = ko.linsolve_bicgstab(Acd, b, maxiter=200, tol=1e-10)
        assert info.converged == 1 and info.numiter == oinfo["numiter"] and info.numops == oinfo["numops"]
        np.testing.assert_allclose(gather(x), ox, rtol=1e-8, atol=1e-10)
        spd = (5.0, -1.0, -1.0, -1.0, -1.0, 0.0, 0.0)
        x, info = kk.linsolve(kk.B200CSR.stencil(ctx, nx, ny, 1, spd), ctx.from_host(b[sl]), None,
                              kk.CG(maxiter=300, tol=1e-10, verbosity=0))
        ox, oinfo = ko.linsolve_cg(ko.stencil_matrix(nx, ny, 1, spd), b, maxiter=300, tol=1e-10)
        assert info.converged == 1 and info.numiter == oinfo["numiter"]
        np.testing.assert_allclose(gather(x), ox, rtol=1e-8, atol=1e-10)

        # Arnoldi eigsolve (complex pair as two sharded device vectors), BlockLanczos, exponentiate
        alg = kk.Arnoldi(orth=kk.cgs2, krylovdim=20, maxiter=4, tol=0.0, verbosity=0)
        vals, vecs, info = kk.eigsolve(opcd, ctx.from_host(x0[sl]), 2, "LR", alg)
        ovals, _, oinfo = ko.eigsolve_arnoldi(Acd, x0, 2, "LR", krylovdim=20, maxiter=4, tol=0.0, orth=ko.Orth(ko.CGS2))
        assert info.numops == oinfo["numops"]
        np.testing.assert_allclose(vals[:2], ovals[:2], rtol=1e-8)
        vg = gather(vecs[0].re) + 1j * gather(vecs[0].im)
        rg = gather(info.residual[0].re) + 1j * gather(info.residual[0].im)
        assert np.linalg.norm(Acd @ vg - vals[0] * vg - rg) < 1e-9
        del vecs, info
        X0 = [ko.splitmix_vector(100 + i, n) for i in range(3)]
        alg = kk.BlockLanczos(krylovdim=18, maxiter=4, tol=0.0, verbosity=0)
        vals, vecs, info = kk.eigsolve(op, kk.Block([ctx.from_host(x[sl]) for x in X0]), 3, "SR", alg)
        ovals, _, oinfo = ko.eigsolve_blocklanczos(A, X0, 3, "SR", krylovdim=18, maxiter=4, tol=0.0)
        assert info.numops == oinfo["numops"] and info.numiter == oinfo["numiter"]
        np.testing.assert_allclose(vals[:3], ovals[:3], rtol=1e-8)
        del vecs, info
        w, info = kk.exponentiate(op, -0.3, ctx.from_host(x0[sl]), kk.Lanczos(orth=kk.cgs2, krylovdim=15, tol=1e-10, verbosity=0))
        ow, oinfo = ko.expintegrator(A, -0.3, (x0,), "lanczos", ko.Orth(ko.CGS2), krylovdim=15, tol=1e-10)
        assert info.converged == 1 and info.numops == oinfo["numops"]
        np.testing.assert_allclose(gather(w), ow, rtol=1e-9, atol=1e-12)
        report["ok"] = True
        ctx.close()

        # the host-buffer (end-to-end) entry in its row-sharded form — what bench.py's e2e leg calls on every rank:
        # local rows as a CSR triple with GLOBAL column indices, the local slice of x0, Ritz vectors into
        # caller-provided host arrays; a context per call, twice (the second call reuses what the first left)
        Aloc = A[sl].tocsr()
        alg = kk.Lanczos(orth=kk.cgs2, krylovdim=20, maxiter=3, tol=0.0, verbosity=0)
        ovals, ovecs, oinfo = ko.eigsolve_lanczos(A, x0, 2, "SR", krylovdim=20, maxiter=3, tol=0.0, orth=ko.Orth(ko.CGS2))
        for _ in range(2):
            outs = [np.empty(shard.n_local) for _ in range(2)]
            vals, vecs_h, info = kk.eigsolve((Aloc.indptr.astype(np.int32), Aloc.indices.astype(np.int32), Aloc.data),
                                             x0[sl].copy(), 2, "SR", alg, out_vectors=outs, shard=shard,
                                             nccl_uid=bytes(128))
            assert info.numops == oinfo["numops"]
            np.testing.assert_allclose(vals[:2], ovals[:2], rtol=1e-10)
            assert vecs_h[0] is outs[0]                          # written in place
            parts = [None] * world
            dist.all_gather_object(parts, outs[0])
            vg = np.concatenate(parts)
            assert abs(abs(vg @ ovecs[0]) - 1.0) < 1e-8          # the oracle's Ritz vector up to sign
    if rank == 0:
        np.save(out, np.array(report["lanczos"]))
    dist.barrier()
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("fused", [False, True], ids=["literal", "fused-steps"])
def test_all_drivers_row_sharded_on_two_ranks(tmp_path, fused):
    import torch.multiprocessing as mp
    from oracle import krylov_oracle as ko
    out = str(tmp_path / "vals.npy")
    port = 31500 + (os.getpid() % 2000) + (1 if fused else 0)
    mp.spawn(_worker, args=(2, port, out, fused), nprocs=2, join=True)
    np.testing.assert_allclose(np.load(out), ko.laplace_eigenvalues(30, 22)[:2], rtol=1e-9)
