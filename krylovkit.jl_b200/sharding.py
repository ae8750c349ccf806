"""Row sharding of vectors / basis / operator across ranks (one process per GPU) — SURVEY §8e.

Host-side arithmetic only: which rows a rank owns, which x entries it must receive from its
neighbours for a banded operator (the halo plan that libb200krylov derives on the device in
`plan_halo`, csrc/spmv.cu), and how the NCCL unique id reaches every rank.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass


@dataclass(frozen=True)
class RowShard:
    rank: int
    world: int
    row_offset: int
    n_local: int
    n_global: int


def shard_grid_lines(nx: int, nlines: int, rank: int, world: int) -> RowShard:
    """Contiguous shard made of whole grid lines (x fastest): rows [y0*nx, y1*nx)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside [0, {world})")
    y0 = (nlines * rank) // world
    y1 = (nlines * (rank + 1)) // world
    return RowShard(rank, world, y0 * nx, (y1 - y0) * nx, nx * nlines)


def halo_plan(shard: RowShard, col_min: int, col_max: int, n_local_all: list[int]):
    """Neighbour exchange sizes for local rows whose global columns span [col_min, col_max].
    Returns (halo_lo, halo_hi, ok): entries needed from rank-1 / rank+1, and whether a pure
    nearest-neighbour exchange suffices (else the library falls back to an all-gather)."""
    lo = max(0, shard.row_offset - col_min)
    hi = max(0, col_max - (shard.row_offset + shard.n_local - 1))
    ok = True
    if lo > 0 and (shard.rank == 0 or lo > n_local_all[shard.rank - 1]):
        ok = False
    if hi > 0 and (shard.rank == shard.world - 1 or hi > n_local_all[shard.rank + 1]):
        ok = False
    return lo, hi, ok


def localize_column(g: int, shard: RowShard, halo_lo: int) -> int:
    """global column -> index into [local x | lo halo | hi halo] (k_localize_cols)."""
    if shard.row_offset <= g < shard.row_offset + shard.n_local:
        return g - shard.row_offset
    if g < shard.row_offset:
        return shard.n_local + (g - (shard.row_offset - halo_lo))
    return shard.n_local + halo_lo + (g - (shard.row_offset + shard.n_local))


def broadcast_nccl_uid(dist, lib, device=None) -> bytes:
    """The 128-byte job-unique token every rank passes to b2k_ctx_create_dist: rank 0 creates it and every rank
    receives it through the torch.distributed process group that launched the job (any backend).  It is an
    ncclUniqueId (b2k_nccl_unique_id) when NCCL is in use; with B2K_NO_NCCL=1 — the NVLink peer window is then
    the only transport, which also allows several ranks on ONE GPU — 128 random bytes do: the token only names
    the node-local rendezvous."""
    import os
    payload = [None]
    if dist.get_rank() == 0:
        if os.environ.get("B2K_NO_NCCL", "") == "1":
            payload[0] = os.urandom(128)
        else:
            buf = C.create_string_buffer(128)
            if lib.b2k_nccl_unique_id(buf) != 0:
                raise RuntimeError("b2k_nccl_unique_id failed")
            payload[0] = bytes(buf.raw)
    dist.broadcast_object_list(payload, src=0)
    return bytes(payload[0])
