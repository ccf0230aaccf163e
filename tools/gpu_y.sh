#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/e2e_chunks.py > gpurun_out/y_e2e.jsonl 2> gpurun_out/y_e2e.err; cat gpurun_out/y_e2e.jsonl; tail -3 gpurun_out/y_e2e.err
KMN_TRACE_RANGES=1 timeout 120 python - <<'PY' 2> gpurun_out/y_trace.txt
import sys, time
sys.path.insert(0, ".")
from kamino_b200 import ffi, synth
b = synth.stage(synth.bacterial(100, seed=synth.SEED0 + 1))
pin = ffi.PinnedBuffer(b.residues.size); pin.array[:] = b.residues
with ffi.Context(13, "sr6") as ctx:
    for i in range(5):
        ctx.reset_table(); ctx.synchronize()
        t0 = time.perf_counter(); ctx.count_batch(pin.array, b.rec_off, b.sp_rec_off); ctx.finalize_table()
        print(f"call {i}: {(time.perf_counter()-t0)*1e3:.3f} ms wall", file=sys.stderr); ctx.release_batches()
PY
tail -32 gpurun_out/y_trace.txt
