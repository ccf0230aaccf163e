#!/bin/bash
# device timeline of kmn_count_batch's copy / frame / compute streams (KMN_TRACE_RANGES) on the bench workload
mkdir -p gpurun_out
KMN_TRACE_RANGES=1 timeout 120 python - <<'PY' 2> gpurun_out/trace.txt
import sys, time
sys.path.insert(0, ".")
from kamino_b200 import ffi, synth
b = synth.stage(synth.bacterial(100, seed=synth.SEED0 + 1))
pin = ffi.PinnedBuffer(b.residues.size); pin.array[:] = b.residues
pr, h_rec = ffi.PinnedBuffer.copy_of(b.rec_off); ps, h_sp = ffi.PinnedBuffer.copy_of(b.sp_rec_off)
with ffi.Context(13, "sr6") as ctx:
    for i in range(5):
        ctx.reset_table(); ctx.synchronize()
        t0 = time.perf_counter(); ctx.count_batch(pin.array, h_rec, h_sp); ctx.finalize_table()
        print(f"call {i}: {(time.perf_counter()-t0)*1e3:.3f} ms wall", file=sys.stderr); ctx.release_batches()
PY
tail -${1:-34} gpurun_out/trace.txt
