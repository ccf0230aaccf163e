#!/usr/bin/env python
"""Probe: do two pass-1 pipelines on two streams of one GPU overlap usefully?  Two contexts, each staged
with half of the 100 proteomes, counted from two host threads (ctypes releases the GIL), against one context
with all 100."""
import sys, threading, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from kamino_b200 import ffi, synth

species = synth.bacterial(100)
full = synth.stage(species)
halves = [synth.stage(species[:50]), synth.stage(species[50:])]
n_iter = 20

with ffi.Context(13, "sr6") as c:
    bid = c.stage_batch(full.residues, full.rec_off, full.sp_rec_off)
    for _ in range(3):
        c.reset_table(); c.count_staged(bid)
    t0 = time.perf_counter()
    for _ in range(n_iter):
        c.reset_table(); c.count_staged(bid)
    t_one = (time.perf_counter() - t0) / n_iter

ctxs = [ffi.Context(13, "sr6") for _ in range(2)]
bids = [c.stage_batch(h.residues, h.rec_off, h.sp_rec_off) for c, h in zip(ctxs, halves)]
def work(i):
    for _ in range(n_iter):
        ctxs[i].reset_table(); ctxs[i].count_staged(bids[i])
for i in range(2):
    for _ in range(3):
        ctxs[i].reset_table(); ctxs[i].count_staged(bids[i])
t0 = time.perf_counter()
th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
[t.start() for t in th]; [t.join() for t in th]
t_two = (time.perf_counter() - t0) / n_iter
t0 = time.perf_counter()
work(0); work(1)
t_seq = (time.perf_counter() - t0) / n_iter
print(f"one context, 100 proteomes: {t_one*1e3:.2f} ms/iter; two contexts x 50 concurrently: {t_two*1e3:.2f} ms/iter; "
      f"the same two sequentially: {t_seq*1e3:.2f} ms/iter")
for c in ctxs: c.close()
