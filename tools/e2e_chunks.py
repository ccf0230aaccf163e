#!/usr/bin/env python
"""kmn_count_batch from pinned host memory (the e2e leg of bench.py) for several copy/compute range schedules
(KMN_H2D_CHUNKS x KMN_H2D_MODE); one JSON line per setting."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    from kamino_b200 import ffi, synth
    b = synth.stage(synth.bacterial(100, seed=synth.SEED0 + 1))
    pin = ffi.PinnedBuffer(b.residues.size)
    pin.array[:] = b.residues
    pin_rec, h_rec = ffi.PinnedBuffer.copy_of(b.rec_off)
    pin_sp, h_sp = ffi.PinnedBuffer.copy_of(b.sp_rec_off)
    settings = [("adaptive", None, None)]
    ratios = os.environ.get("E2E_RATIOS", "1.0,1.5,2.0,3.0").split(",")
    chunks = [int(x) for x in os.environ.get("E2E_CHUNKS", "3,4,5,8").split(",")]
    settings += [("ratio", r, n) for r in ratios for n in chunks]
    for mode, ratio, chunks in settings:
        for _ in (0,):
            for k in ("KMN_H2D_CHUNKS", "KMN_H2D_RATIO", "KMN_H2D_MODE"):
                os.environ.pop(k, None)
            if ratio is not None:
                os.environ["KMN_H2D_RATIO"] = ratio
                os.environ["KMN_H2D_CHUNKS"] = str(chunks)
            with ffi.Context(13, "sr6") as ctx:
                ts = []
                for i in range(9):
                    ctx.reset_table()
                    ctx.synchronize()
                    t0 = time.perf_counter()
                    ctx.count_batch(pin.array, h_rec, h_sp)
                    ctx.finalize_table()
                    ts.append((time.perf_counter() - t0) * 1e3)
                    ctx.release_batches()
                print(json.dumps({"mode": mode, "ratio": ratio, "chunks": chunks, "ms": round(float(np.median(ts[3:])), 3)}), flush=True)
    pin.free()


if __name__ == "__main__":
    main()
