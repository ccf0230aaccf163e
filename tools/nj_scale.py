"""--nj pair counts sharded by row blocks over the ranks of a torchrun job.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/nj_scale.py [--n 5000] [--l 100000] [--reps 3]

Every rank holds the whole synthetic alignment, computes its row block with kmn_pair_counts_rows and the
u32 count blocks are gathered on rank 0.  Prints one JSON line: kernel time (CUDA events on each rank's
stream, max over ranks), the time of the whole sharded call including the gather, and a parity check of a
corner of the matrix against the oracle.
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=5000)
    ap.add_argument("--l", type=int, default=100_000)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from kamino_b200 import ffi, sharding, synth

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rng = np.random.default_rng(1)
    n, L = args.n, args.l
    base = rng.choice(20, size=L, p=synth.LG_PI).astype(np.uint8)
    m = np.tile(base, (n, 1))
    mut = rng.random((n, L)) < 0.05
    m[mut] = rng.integers(0, 20, int(mut.sum()), dtype=np.uint8)
    m[rng.random((n, L)) < 0.05] = 20
    blocks = sharding.pair_row_blocks(n, world)
    with ffi.Context(13, "sr6", device=local) as ctx:
        ctx.pair_counts(m[:64])
        k_ms, call_ms = [], []
        got = None
        for _ in range(args.reps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            got = sharding.pair_counts_sharded(ctx, m)
            call_ms.append((time.perf_counter() - t0) * 1e3)
            t = torch.tensor([ctx.stage_ms("pair_counts")], device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            k_ms.append(float(t.item()))
    if rank == 0:
        import oracle_ffi
        n_cpu = 40
        wv, wm = oracle_ffi.load().pair_counts(m[:n_cpu], threads=8)
        iu = np.triu_indices(n_cpu, 1)
        same = bool(np.array_equal(got[0][:n_cpu, :n_cpu][iu], wv[iu]) and np.array_equal(got[1][:n_cpu, :n_cpu][iu], wm[iu]))
        # the last rows live on the last rank: compare them with a direct recount
        tail = m[n - 8:].astype(np.int16)
        ok = (tail[:, None, :] < 20) & (tail[None, :, :] < 20)
        tv = ok.sum(-1).astype(np.uint32)
        tm = (ok & (tail[:, None, :] != tail[None, :, :])).sum(-1).astype(np.uint32)
        it = np.triu_indices(8, 1)
        same = same and bool(np.array_equal(got[0][n - 8:, n - 8:][it], tv[it]) and np.array_equal(got[1][n - 8:, n - 8:][it], tm[it]))
        kern = float(np.median(k_ms))
        pairs = n * (n - 1) // 2
        print(json.dumps({"what": "pair_counts_sharded", "n_gpus": world, "n": n, "L": L, "row_blocks": blocks,
                          "kernel_ms_max_over_ranks": kern, "pair_columns_per_s": pairs * L / (kern * 1e-3),
                          "call_ms_with_gather": float(np.median(call_ms)), "corners_equal_cpu": same}))
        assert same
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
