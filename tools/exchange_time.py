#!/usr/bin/env python
"""Per-rank timing of the table exchange (torchrun, one rank per GPU): device time of kmn_exchange_reduce
(stage 'finalize': split + exchange kernel + join, including the in-kernel wait for the slowest peer) next to
pass 1, for both exchange flavours.  One JSON line per rank."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    import torch
    import torch.distributed as dist
    from kamino_b200 import ffi, sharding, synth
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_sp = int(os.environ.get("KMN_BENCH_SPECIES", "100"))
    b = synth.stage(synth.bacterial(n_sp, seed=synth.SEED0 + 1 + 1000 * rank))
    out = {"rank": rank, "world": world, "gap": os.environ.get("XT_GAP", "none"), "noabort": os.environ.get("KMN_EXCHANGE_NOABORT", "0")}
    for planes in ("1", "0"):
        os.environ["KMN_EXCHANGE_PLANES"] = planes
        with ffi.Context(13, "sr6", device=local) as ctx:
            assert sharding.setup_fused_exchange(ctx)
            bid = ctx.stage_batch(b.residues, b.rec_off, b.sp_rec_off)
            p1, ex, wall, only = [], [], [], []
            gap = os.environ.get("XT_GAP", "none")
            for i in range(8):
                dist.barrier(); torch.cuda.synchronize()
                t0 = time.perf_counter()
                ctx.reset_table()
                ctx.count_staged(bid, b.n_sp)
                s = sum(ctx.stage_ms(x) for x in ("frame", "bloom", "cms"))
                if gap == "sleep":
                    time.sleep(0.0005)
                elif gap == "barrier":
                    dist.barrier(); torch.cuda.synchronize()
                ctx.exchange_reduce()
                dt = (time.perf_counter() - t0) * 1e3
                if i >= 3:
                    p1.append(s); ex.append(ctx.stage_ms("finalize")); wall.append(dt)
            for i in range(6):   # exchange alone, ranks aligned by a barrier right before it
                ctx.reset_table()
                ctx.count_staged(bid, b.n_sp)
                dist.barrier(); torch.cuda.synchronize()
                t0 = time.perf_counter()
                ctx.exchange_reduce()
                dt = (time.perf_counter() - t0) * 1e3
                if i >= 2:
                    only.append((ctx.stage_ms("finalize"), dt))
            out[f"planes={planes}"] = {"pass1_ms": round(float(np.median(p1)), 3), "exchange_ms": round(float(np.median(ex)), 3),
                                       "wall_ms": round(float(np.median(wall)), 3),
                                       "exchange_alone_device_ms": round(float(np.median([x[0] for x in only])), 3),
                                       "exchange_alone_wall_ms": round(float(np.median([x[1] for x in only])), 3)}
            dist.barrier()
    print(json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
