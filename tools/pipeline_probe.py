#!/usr/bin/env python
"""Where a step of the pipelined multi-batch job (kmn_stage_batch_async + kmn_count_staged + kmn_release_batch) spends
its time: host wall time of each call and the device time of the count's stages, against the same count with nothing
being copied.  One JSON line."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    from kamino_b200 import ffi, synth
    b = synth.stage(synth.bacterial(100, seed=synth.SEED0 + 1))
    pins = [ffi.PinnedBuffer.copy_of(a) for a in (b.residues, b.rec_off, b.sp_rec_off)]
    h = [v for _, v in pins]
    n = 12
    with ffi.Context(13, "sr6") as ctx:
        # reference: resident batch, nothing in flight
        bid = ctx.stage_batch(*h)
        ctx.reset_table()
        alone = []
        for i in range(6):
            t0 = time.perf_counter()
            ctx.count_staged(bid, b.n_sp)
            alone.append(((time.perf_counter() - t0) * 1e3, sum(ctx.stage_ms(s) for s in ("frame", "bloom", "cms"))))
        ctx.release_batches()
        ctx.reset_table()
        rows = []
        t_job = time.perf_counter()
        nxt = ctx.stage_batch_async(*h)
        for i in range(n):
            cur = nxt
            t0 = time.perf_counter()
            if i + 1 < n:
                nxt = ctx.stage_batch_async(*h)
            t1 = time.perf_counter()
            ctx.count_staged(cur, b.n_sp)
            t2 = time.perf_counter()
            dev = sum(ctx.stage_ms(s) for s in ("frame", "bloom", "cms"))
            ctx.release_batch(cur)
            t3 = time.perf_counter()
            rows.append([(t1 - t0) * 1e3, (t2 - t1) * 1e3, dev, (t3 - t2) * 1e3])
        ctx.finalize_table()
        job_ms = (time.perf_counter() - t_job) * 1e3
    r = np.array(rows[2:-1])
    a = np.array(alone[2:])
    print(json.dumps({"resident_count_wall_ms": round(float(np.median(a[:, 0])), 3), "resident_count_device_ms": round(float(np.median(a[:, 1])), 3),
                      "pipelined": {"stage_async_wall_ms": round(float(np.median(r[:, 0])), 3), "count_staged_wall_ms": round(float(np.median(r[:, 1])), 3),
                                    "count_staged_device_ms": round(float(np.median(r[:, 2])), 3), "release_wall_ms": round(float(np.median(r[:, 3])), 3),
                                    "job_ms_per_step": round(job_ms / n, 3)}}))
    for p, _ in pins:
        p.free()


if __name__ == "__main__":
    main()
