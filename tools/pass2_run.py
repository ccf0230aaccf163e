#!/usr/bin/env python
"""Pass 2 on the bench workload (100 proteomes): count + finalize, then N x (scan, anchor pairing, grouping); one JSON line
with the library's stage times.  Meant to run under ncu (launch list / --set full of the pass-2 kernels)."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    from kamino_b200 import ffi, synth
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    b = synth.stage(synth.bacterial(100, seed=synth.SEED0 + 1))
    min_needed = int(np.ceil(np.float32(0.85) * np.float32(b.n_sp)))
    with ffi.Context(13, "sr6") as ctx:
        bid = ctx.stage_batch(b.residues, b.rec_off, b.sp_rec_off)
        ctx.count_staged(bid, b.n_sp)
        ctx.finalize_table()
        scan, pair, grp = [], [], []
        for _ in range(reps):
            ns, ne = ctx.scan_batch(bid, min_needed)
            scan.append(ctx.stage_ms("scan"))
            n_pairs = ctx.pair_batch(bid, 35)
            pair.append(ctx.stage_ms("pair_batch"))
            n_obs, n_groups, n_valid = ctx.group_batches([bid], min_needed)
            grp.append(ctx.stage_ms("group_observations"))
    print(json.dumps({"residues": b.n_residues, "min_needed": min_needed, "scan_ms": scan, "pair_batch_ms": pair, "group_ms": grp,
                      "starts": int(ns), "ends": int(ne), "anchor_pairs": int(n_pairs), "groups": int(n_groups),
                      "groups_passing_length_selection": int(n_valid)}))


if __name__ == "__main__":
    main()
