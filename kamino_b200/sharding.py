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


def clamp_u16(acc: np.ndarray) -> np.ndarray:
    """Host restatement of the finalize kernel (tests only)."""
    return np.minimum(acc.astype(np.int64) & 0xFFFFFFFF, 65535).astype(np.uint16)
