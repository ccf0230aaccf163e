#!/usr/bin/env python
"""Per-kernel device times of pass 1 on the bench workload (100 proteomes, inputs resident in HBM), median of N
steps, one JSON line.  For A/B runs of tuning builds:  KAMINO_B200_LIB=<lib> python tools/stage_times.py"""
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    from kamino_b200 import ffi, synth
    n_sp = int(os.environ.get("KMN_BENCH_SPECIES", "100"))
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    b = synth.stage(synth.bacterial(n_sp, seed=synth.SEED0 + 1))
    names = ["frame", "bloom_partition_kernel", "bloom_bucket_kernel", "bloom", "cms_partition_kernel", "cms_apply_kernel", "cms"]
    with ffi.Context(13, "sr6") as ctx:
        bid = ctx.stage_batch(b.residues, b.rec_off, b.sp_rec_off)
        acc = {n: [] for n in names}
        new = None
        accumulate = os.environ.get("KMN_STAGE_ACCUMULATE") == "1"   # batches 2.. of a job: cms_apply merges into a live table
        for i in range(3 + steps):
            if not (accumulate and i > 0):
                ctx.reset_table()
            new = ctx.count_staged(bid, b.n_sp)
            if not accumulate:
                ctx.finalize_table()
            if i >= 3:
                for n in names:
                    acc[n].append(ctx.stage_ms(n))
        if accumulate:
            ctx.finalize_table()
        rows, h = ctx.table_checksum()
    med = {n: round(float(np.median(v)), 4) for n, v in acc.items()}
    med["pass1_sum"] = round(med["frame"] + med["bloom"] + med["cms"], 4)
    print(json.dumps({"lib": os.environ.get("KAMINO_B200_LIB", "default"), "species": n_sp, "residues": b.n_residues,
                      "ms": med, "new_kmers": int(new.sum()), "table_hash": f"{h:016x}"}))


if __name__ == "__main__":
    main()
