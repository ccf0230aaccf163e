"""Multi-GPU plumbing for the table build (one process per GPU, torch.distributed).

Species shard across ranks (a species never spans ranks, so the per-species Bloom stays rank-local);
the count-min sketch is LINEAR, so one exchange step — the sum of the per-rank u32 accumulators —
followed by the local clamp gives every rank the table the single-process reference would build:
    final = min(sum_over_ranks(acc), 65535)          (reference: src/proba_filter.rs:146-177)
"""
from __future__ import annotations

import numpy as np


def shard_species(n_species: int, world: int, rank: int) -> np.ndarray:
    """Round-robin assignment (north_star: 'proteome batches are distributed round-robin')."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    return np.arange(rank, n_species, world, dtype=np.int64)


def select_species(residues, rec_off, sp_rec_off, species: np.ndarray):
    """Sub-batch holding only `species` (in the given order) of a staged batch."""
    rec_off = np.asarray(rec_off, dtype=np.uint64)
    sp_rec_off = np.asarray(sp_rec_off, dtype=np.uint64)
    parts, lens, sp = [], [], [0]
    for s in species:
        r0, r1 = int(sp_rec_off[s]), int(sp_rec_off[s + 1])
        b0, b1 = int(rec_off[r0]), int(rec_off[r1])
        parts.append(residues[b0:b1])
        lens.append(np.diff(rec_off[r0:r1 + 1].astype(np.int64)))
        sp.append(sp[-1] + (r1 - r0))
    out_res = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
    all_lens = np.concatenate(lens) if lens else np.zeros(0, np.int64)
    out_rec = np.zeros(len(all_lens) + 1, dtype=np.uint64)
    np.cumsum(all_lens, out=out_rec[1:])
    return out_res, out_rec, np.asarray(sp, dtype=np.uint64)


def exchange_accumulators(acc, group=None):
    """In-place sum of the per-rank accumulators (torch tensor, int32 view of the u32 cells)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return acc


def setup_fused_exchange(ctx, group=None) -> bool:
    """All-gather the CUDA IPC handles and map the peers' tables (once per context).  Returns False
    when peer mapping is unavailable (the caller then uses exchange_accumulators + finalize)."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    handles = [None] * world
    dist.all_gather_object(handles, ctx.exchange_export(), group=group)
    ok = True
    try:
        ctx.exchange_setup(world, rank, [h[0] for h in handles], [h[1] for h in handles])
    except Exception:
        ok = False
    flags = [None] * world
    dist.all_gather_object(flags, ok, group=group)
    return all(flags)


def fused_exchange(ctx, group=None):
    """reduce-scatter + snapshot + all-gather in one kernel per rank over NVLink peer memory.  The ranks
    synchronise inside the kernel through signal words in peer memory (no host barrier): when the call
    returns, every peer has written its slice into this rank's table and nobody still reads this rank's."""
    ctx.exchange_reduce()          # stream-ordered after the counting kernels; synchronizes before returning


PAIR_TILE_ROWS = 64   # row granularity of kmn_pair_counts_rows


def pair_row_blocks(n: int, world: int) -> list[tuple[int, int]]:
    """Contiguous row blocks [begin, end) of the n x n pair matrix, one per rank, aligned to the kernel's
    64-row tiles and balanced by the number of upper-triangle tiles (row tile t holds n_tiles - t of them).
    Replaces rayon's split of the row loop of compute_distance_matrix (src/phylo.rs:124-138)."""
    if world < 1:
        raise ValueError("world must be >= 1")
    n_tiles = (n + PAIR_TILE_ROWS - 1) // PAIR_TILE_ROWS
    total = n_tiles * (n_tiles + 1) // 2
    blocks, t, done = [], 0, 0
    for r in range(world):
        target = total * (r + 1) // world      # tiles that ranks 0..r should cover together
        t0 = t
        # take row tiles while that brings the running total closer to the target
        while t < n_tiles and (r == world - 1 or abs(done + (n_tiles - t) - target) <= abs(done - target)):
            done += n_tiles - t
            t += 1
        blocks.append((min(t0 * PAIR_TILE_ROWS, n), min(t * PAIR_TILE_ROWS, n)))
    return blocks


class _DevArray:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}


def pair_counts_sharded(worker, mapped: np.ndarray, group=None, dst: int = 0):
    """--nj counts with the rows of the pair matrix sharded over the ranks of `group`.

    worker is a kamino_b200.ffi.Context (the counts of each rank's row block stay on its GPU and are
    gathered device to device) or, for host-side tests, a callable
    count_rows(mapped, row_begin, row_end) -> (valid, mismatch) of shape (rows, n).
    Every rank holds the whole alignment (small: n x L bytes); the u32 count blocks are gathered on
    `dst`, which returns the full (n, n) matrices — the other ranks return None.  This gather is the
    only collective."""
    import torch
    import torch.distributed as dist
    n = mapped.shape[0]
    on_device = hasattr(worker, "pair_counts_device")
    count_rows = worker.pair_counts_rows if on_device else worker
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return count_rows(mapped, 0, n)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    blocks = pair_row_blocks(n, world)
    b0, b1 = blocks[rank]
    max_rows = max(e - b for b, e in blocks)
    if on_device:
        dev = torch.device("cuda", torch.cuda.current_device())
        count_rows(mapped, b0, b1, fetch=False)
        mine = torch.zeros((2, max_rows, n), dtype=torch.int32, device=dev)   # int32 view of the u32 counts
        if b1 > b0:
            pv, pm, rows, cols = worker.pair_counts_device()
            assert (rows, cols) == (b1 - b0, n)
            mine[0, :rows] = torch.as_tensor(_DevArray(pv, (rows, n), "<i4"), device=dev)
            mine[1, :rows] = torch.as_tensor(_DevArray(pm, (rows, n), "<i4"), device=dev)
    else:
        dev = torch.device("cpu")
        valid, mism = count_rows(mapped, b0, b1)
        mine = torch.zeros((2, max_rows, n), dtype=torch.int32)
        if b1 > b0:
            mine[0, : b1 - b0] = torch.from_numpy(valid.view(np.int32))
            mine[1, : b1 - b0] = torch.from_numpy(mism.view(np.int32))
    parts = [torch.empty_like(mine) for _ in range(world)] if rank == dst else None
    dist.gather(mine, parts, dst=dst, group=group)
    if rank != dst:
        return None
    full_v = np.empty((n, n), dtype=np.uint32)
    full_m = np.empty((n, n), dtype=np.uint32)
    tv, tm = torch.from_numpy(full_v.view(np.int32)), torch.from_numpy(full_m.view(np.int32))
    for (e0, e1), part in zip(blocks, parts):
        if e1 > e0:
            tv[e0:e1].copy_(part[0, : e1 - e0])
            tm[e0:e1].copy_(part[1, : e1 - e0])
    return full_v, full_m


def clamp_u16(acc: np.ndarray) -> np.ndarray:
    """Host restatement of the finalize kernel (tests only)."""
    return np.minimum(acc.astype(np.int64) & 0xFFFFFFFF, 65535).astype(np.uint16)
