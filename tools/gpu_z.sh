#!/bin/bash
# 1000-proteome job from pinned host memory: long range schedule (default) vs the 5-range one (KMN_H2D_CHUNKS=5)
mkdir -p gpurun_out
timeout 600 python - <<'PY' > gpurun_out/z_long.jsonl 2> gpurun_out/z_long.err
import sys, time, os, json
sys.path.insert(0, ".")
import numpy as np
from kamino_b200 import ffi, synth
b = synth.stage(synth.bacterial(1000, seed=synth.SEED0 + 2))
pin = ffi.PinnedBuffer(b.residues.size); pin.array[:] = b.residues
for setting in (None, "5", "16", None):
    os.environ.pop("KMN_H2D_CHUNKS", None)
    if setting: os.environ["KMN_H2D_CHUNKS"] = setting
    with ffi.Context(13, "sr6") as ctx:
        ts = []; hs = None
        for i in range(6):
            ctx.reset_table(); ctx.synchronize()
            t0 = time.perf_counter(); _, new = ctx.count_batch(pin.array, b.rec_off, b.sp_rec_off); ctx.finalize_table()
            ts.append((time.perf_counter()-t0)*1e3)
            if i == 0: rows, hs = ctx.table_checksum(); tot = int(new.sum())
            ctx.release_batches()
        print(json.dumps({"chunks": setting or "long", "ms": round(float(np.median(ts[2:])),3), "hash": f"{hs:016x}", "rows_ok": [int(x) for x in rows] == [tot]*4}), flush=True)
PY
cat gpurun_out/z_long.jsonl; tail -5 gpurun_out/z_long.err
