#!/bin/bash
mkdir -p gpurun_out
KMN_TRACE_RANGES=1 python - > gpurun_out/q_trace.txt 2>&1 <<'PY'
import sys, time
sys.path.insert(0, '.')
import numpy as np
from kamino_b200 import ffi, synth
b = synth.stage(synth.bacterial(100, seed=synth.SEED0 + 1))
pin = ffi.PinnedBuffer(b.residues.size); pin.array[:] = b.residues
with ffi.Context(13, "sr6") as ctx:
    for i in range(6):
        ctx.reset_table(); ctx.synchronize()
        if i == 5: print("=== traced call ===", flush=True); sys.stderr.flush()
        t0 = time.perf_counter()
        ctx.count_batch(pin.array, b.rec_off, b.sp_rec_off)
        dt = (time.perf_counter() - t0) * 1e3
        ctx.finalize_table(); ctx.release_batches()
        print(f"call {i}: {dt:.3f} ms wall", flush=True)
PY
grep -A40 "traced call" gpurun_out/q_trace.txt | head -60
