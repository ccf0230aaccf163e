#!/usr/bin/env python
"""Robustness / throughput at BASELINE.json configs[2] size on ONE GPU: N synthetic bacterial proteomes in a
single batch through pass 1, pass 2, device pairing and grouping.  The table is checked through
size-independent properties (batch-split invariance against two half batches; new-k-mer counts)."""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--species", type=int, default=1000)
    args = ap.parse_args()
    from kamino_b200 import ffi, synth
    t0 = time.time()
    species = synth.bacterial(args.species)
    b = synth.stage(species)
    gen_s = time.time() - t0
    half = args.species // 2
    ha, hb = synth.stage(species[:half]), synth.stage(species[half:])
    min_needed = int(np.ceil(np.float32(0.85) * np.float32(args.species)))
    with ffi.Context(13, "sr6") as ctx:
        t0 = time.time()
        bid, new = ctx.count_batch(b.residues, b.rec_off, b.sp_rec_off)
        ctx.finalize_table()
        p1_s = time.time() - t0
        table = ctx.table_snapshot()
        t0 = time.time()
        ns, ne = ctx.scan_batch(bid, min_needed)
        scan_ms = ctx.stage_ms("scan")
        n_pairs = ctx.pair_batch(bid, 35)
        pair_ms = ctx.stage_ms("pair_batch")
        p2_s = time.time() - t0
        ctx.release_batches()
        ctx.reset_table()
        _, na = ctx.count_batch(ha.residues, ha.rec_off, ha.sp_rec_off)
        _, nb = ctx.count_batch(hb.residues, hb.rec_off, hb.sp_rec_off)
        ctx.finalize_table()
        same = bool(np.array_equal(ctx.table_snapshot(), table)) and bool(np.array_equal(np.concatenate([na, nb]), new))
    print(json.dumps({"what": "scale_check", "species": b.n_sp, "residues": b.n_residues, "generate_s": gen_s,
                      "pass1_wall_s_host_buffers": p1_s, "pass1_residues_per_s": b.n_residues / p1_s,
                      "passing_fraction": float(new.sum()) / b.n_residues, "table_max": int(table.max()),
                      "scan_ms": scan_ms, "pair_batch_ms": pair_ms, "pass2_wall_s": p2_s, "starts": ns, "ends": ne,
                      "anchor_pairs": n_pairs, "split_invariant": same}))
    assert same


if __name__ == "__main__":
    main()
