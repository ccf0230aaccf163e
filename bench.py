#!/usr/bin/env python
"""bench.py — k-mer table build (pass 1) throughput on B200, BASELINE.json's metric.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference            # CPU arm: C++ restatement of kamino 2.0.0

A "step" is one pass of the hot path over one batch: 100 synthetic bacterial proteomes per GPU
(BASELINE.json configs[1]) -> recode/frame, exact per-species Bloom, count-min build into the
saturating u16 table (+ at N>1 the table exchange: fused NVLink peer kernel, or NCCL sum).
  value : residues/s, inputs already resident in HBM (kmn_count_staged + kmn_finalize_table)
  e2e   : residues/s through the C ABI from pinned HOST buffers (kmn_count_batch: H2D inside the
          timed region) + finalize + D2H of the per-species new-k-mer counts
The CPU oracle is only ever used here for the cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "residues/sec k-mer table build"
UNIT = "residues/s"
K, SCHEME_NAME, SCHEME = 13, "sr6", 1
N_SPECIES = int(os.environ.get("KMN_BENCH_SPECIES", "100"))   # proteomes per GPU (weak scaling); 100 = BASELINE configs[1]


# stdout carries the ONE JSON line and nothing else: libraries that chat on fd 1 (NCCL prints its version
# banner there) are pointed at stderr for the whole run, and emit() writes to the saved descriptor.
_JSON_FD = None


def claim_stdout():
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, line)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.samples, self.proc = index, [], None
        self.windows, self._open = [], None

    def start(self):
        """Started BEFORE the warm-up: nvidia-smi needs ~0.1 s to come up, the timed region can be shorter."""
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.monotonic(), line.strip()))

    def wait_first(self, timeout: float = 3.0):
        t_end = time.monotonic() + timeout
        while self.proc and not self.samples and time.monotonic() < t_end:
            time.sleep(0.005)

    def begin(self):
        self._open = time.monotonic()

    def end(self):
        if self._open is not None:
            self.windows.append((self._open, time.monotonic()))
            self._open = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def digest(lines):
            sm, mx, reasons = [], [], set()
            for s in lines:
                f = [x.strip() for x in s.split(",")]
                if len(f) < 6:
                    continue
                try:
                    sm.append(float(f[0])); mx.append(float(f[1]))
                except ValueError:
                    continue
                for nme, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            return sm, mx, reasons

        timed = [l for t, l in self.samples if any(a <= t <= b for a, b in self.windows)]
        sm, mx, reasons = digest(timed)
        window = "timed regions"
        if len(sm) < 2 and self.windows:   # regions shorter than the sampling period: everything from the first timed step on
            sm, mx, reasons = digest([l for t, l in self.samples if t >= self.windows[0][0]])
            window = "from the first timed step to the end of the run"
        if not sm:
            sm, mx, reasons = digest([l for _, l in self.samples])
            window = "whole run under load (warm-up + timed)"
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


def make_batch(rank: int):
    from kamino_b200 import synth
    species = synth.bacterial(N_SPECIES, seed=synth.SEED0 + 1 + 1000 * rank)
    return synth.stage(species)


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_pass1(batch, threads: int, n_species: int | None = None):
    """Oracle pass 1 (C++ restatement of kamino 2.0.0, species-parallel) on the host cores."""
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_ffi
    orc = oracle_ffi.load()
    n_sp = batch.n_sp if n_species is None else min(n_species, batch.n_sp)
    rec_end = int(batch.sp_rec_off[n_sp])
    byte_end = int(batch.rec_off[rec_end])
    residues = batch.residues[:byte_end]
    rec_off, sp_rec_off = batch.rec_off[:rec_end + 1], batch.sp_rec_off[:n_sp + 1]
    n_res = int(np.count_nonzero(residues != 10))
    h = orc.cmsa_new()
    try:
        t0 = time.perf_counter()
        new = orc.count_batch(h, residues, rec_off, sp_rec_off, K, SCHEME, threads=threads)
        snap = np.array(orc.cmsa_table(h))  # snapshot copy (proba_filter.rs:171-177)
        dt = time.perf_counter() - t0
        del snap
    finally:
        orc.cmsa_free(h)
    return n_res, dt, int(new.sum()), n_sp


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    batch = make_batch(0)
    threads = host_threads()
    # the step is the same workload as the GPU arm (100 proteomes) unless that would take more than ~10 s of
    # CPU time per step; then a bounded sample of whole species
    n_res, dt, _, n_sp = cpu_pass1(batch, threads, n_species=2 * threads)
    per_species = dt / n_sp
    n_sample = N_SPECIES if per_species * N_SPECIES <= 10.0 else int(max(threads, min(N_SPECIES, 10.0 / max(per_species, 1e-6))))
    for _ in range(args.warmup):
        cpu_pass1(batch, threads, n_species=n_sample)
    tot_res, tot_dt = 0, 0.0
    for _ in range(args.steps):
        n_res, dt, _, _ = cpu_pass1(batch, threads, n_species=n_sample)
        tot_res += n_res
        tot_dt += dt
    value = tot_res / tot_dt
    sample = f"{n_sample} of {N_SPECIES} species per step ({tot_res // args.steps} residues), pass 1 + snapshot"
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"{N_SPECIES} synthetic bacterial proteomes (~4k proteins x ~330 aa), k={K}, {SCHEME_NAME}",
                   "note": "C++ restatement of kamino 2.0.0 (not the Rust binary: no cargo/rustc in the image)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })
    return 0


class _DevArray:
    """Expose a raw device pointer to torch through __cuda_array_interface__ (plumbing only)."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from kamino_b200 import ffi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the k-mer table build has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t0 = time.time()
    batch = make_batch(rank)
    log(f"[rank {rank}] synthetic batch: {batch.n_sp} species, {batch.n_residues} residues, "
        f"{batch.residues.size} bytes staged ({time.time() - t0:.1f}s)")
    n_res_total = batch.n_residues
    if world > 1:   # every rank stages its own proteomes: the job's residue count is the sum over ranks
        t = torch.tensor([batch.n_residues], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        n_res_total = int(t.item())

    ctx = ffi.Context(K, SCHEME, device=local_rank)
    acc_t = None
    if world > 1:
        acc_t = torch.as_tensor(_DevArray(ctx.accum_device_ptr(), ffi.CMS_CELLS, "<i4"), device=f"cuda:{local_rank}")

    # pinned host copies for the end-to-end leg
    pin_res = torch.empty(batch.residues.size, dtype=torch.uint8).pin_memory()
    pin_res.numpy()[:] = batch.residues
    pin_rec = torch.from_numpy(batch.rec_off.astype(np.int64)).pin_memory()
    pin_sp = torch.from_numpy(batch.sp_rec_off.astype(np.int64)).pin_memory()
    h_res, h_rec, h_sp = pin_res.numpy(), pin_rec.numpy().view(np.uint64), pin_sp.numpy().view(np.uint64)

    bid = ctx.stage_batch(h_res, h_rec, h_sp)

    from kamino_b200 import sharding

    fused = False
    if world > 1 and os.environ.get("KMN_EXCHANGE", "fused") == "fused":
        fused = sharding.setup_fused_exchange(ctx)       # CUDA IPC peer mapping over NVLink
        log(f"[rank {rank}] fused NVLink exchange: {'on' if fused else 'unavailable, using NCCL all-reduce'}")

    def exchange_and_finalize():
        if world == 1:
            ctx.finalize_table()
        elif fused:
            sharding.fused_exchange(ctx)                 # one kernel: P2P reduce + clamp + all-gather
        else:
            ctx.table_widen()                            # u16 cells -> summable u32 copy
            sharding.exchange_accumulators(acc_t)        # the sketch is linear: sum of the per-rank tables
            torch.cuda.synchronize()
            ctx.finalize_table()

    def step_resident():
        ctx.reset_table()
        new = ctx.count_staged(bid, batch.n_sp)
        exchange_and_finalize()
        return new

    def step_e2e():
        ctx.reset_table()
        _, new = ctx.count_batch(h_res, h_rec, h_sp)   # H2D of residues + offsets inside
        exchange_and_finalize()
        return new                                       # D2H: per-species counts

    sampler = ClockSampler(local_rank)
    sampler.start()                      # comes up during the warm-up; only samples inside the timed regions count
    for _ in range(args.warmup):
        new = step_resident()
    u = float(new.sum()) / batch.n_residues
    sampler.wait_first()

    # ---- timed: resident inputs
    stage_acc = {"frame": 0.0, "bloom": 0.0, "cms": 0.0, "finalize": 0.0, "cms_partition_kernel": 0.0,
                 "bloom_bucket_kernel": 0.0}
    launches0 = ctx.launch_count()
    lib_stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local_rank))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.begin()
    ev0.record(lib_stream)
    for _ in range(args.steps):
        step_resident()
        for s in stage_acc:
            stage_acc[s] += ctx.stage_ms(s)
    ev1.record(lib_stream)
    barrier()
    sampler.end()
    elapsed = ev0.elapsed_time(ev1) * 1e-3     # device clock, CUDA events on the launching stream
    launches = ctx.launch_count() - launches0
    if world > 1:
        t = torch.tensor([elapsed], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t.item())
    value = n_res_total * args.steps / elapsed

    # ---- timed: end to end from pinned host buffers
    ctx.release_batches()
    e2e_steps = 0 if args.profile else args.steps
    if e2e_steps:
        step_e2e(); ctx.release_batches()
    barrier()
    sampler.begin()
    ev0.record(lib_stream)
    for _ in range(e2e_steps):
        step_e2e()
        ctx.release_batches()
    ev1.record(lib_stream)
    barrier()
    sampler.end()
    clocks = sampler.stop()
    e2e_elapsed = max(ev0.elapsed_time(ev1) * 1e-3, 1e-9)
    if world > 1:
        t = torch.tensor([e2e_elapsed], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_elapsed = float(t.item())
    e2e_value = n_res_total * args.steps / e2e_elapsed
    h2d = int(batch.residues.size + batch.rec_off.nbytes + batch.sp_rec_off.nbytes)
    d2h = int(batch.n_sp * 8 + (batch.n_sp + 1) * 4)

    # ---- roofline of the dominant kernel (device time from CUDA events on the library's stream)
    # cms_partition_kernel (one launch per step) is the longest kernel of pass 1: it hashes every k-mer that
    # passed the Bloom filter into its 4 count-min cells and splits them over 4096 shared-memory-sized
    # buckets.  SURVEY.md 8d: pass 1 moves B1 = 1 + 16u algorithmic bytes per residue (1 residue byte +
    # 4 cells x (2 B read + 2 B write) per passing k-mer); the launch carries all of them for the batch.
    peak, peak_src = measured_peaks()
    stage_ms = {s: v / args.steps for s, v in stage_acc.items()}
    alg_bytes = (1.0 + 16.0 * u) * batch.n_residues
    dom_ms = stage_ms["cms_partition_kernel"]
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
    pass1_ms = stage_ms["frame"] + stage_ms["bloom"] + stage_ms["cms"] + stage_ms["finalize"]
    traffic = None
    tp = ROOT / "profiles" / "roofline_traffic.json"
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text())["cms_partition_kernel"]["dram_bytes_per_launch"]
        except Exception:
            traffic = None
    roofline = {
        "bound": "hbm", "kernel": "cms_partition_kernel", "launches_per_step": 1, "achieved": achieved,
        "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_residue": 1.0 + 16.0 * u, "algorithmic_bytes_per_launch": alg_bytes,
        "new_kmer_fraction_u": u, "kernel_ms_per_launch": dom_ms, "stage_ms": stage_ms,
        "whole_pass1_frac": alg_bytes / (pass1_ms * 1e-3) / 1e9 / peak,
        "note": "integer path: every pass-1 kernel is bound by instruction issue / shared-memory atomics, not by "
                "HBM bytes (ncu: issue slots 65-77 % busy, DRAM < 10 %); see DESIGN.md 4",
    }

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.profile:
        threads = host_threads()
        n_res, dt, _, n_sp = cpu_pass1(batch, threads, n_species=8)
        n_sample = int(max(8, min(N_SPECIES, 15.0 / max(dt / n_sp, 1e-6))))
        n_res, dt, _, n_sp = cpu_pass1(batch, threads, n_species=n_sample)
        cpu_baseline = {"value": n_res / dt, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"{n_sp} of {N_SPECIES} species ({n_res} residues), pass 1 + snapshot, "
                                  "C++ restatement of kamino 2.0.0"}

    if rank == 0:
        emit({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"{N_SPECIES} synthetic bacterial proteomes (~4k proteins x ~330 aa) per GPU, k={K}, {SCHEME_NAME}",
                       "residues_per_gpu": batch.n_residues, "species_per_gpu": batch.n_sp,
                       "l2": "per step the kernels stream 122 MB of residues, ~2 GB of partition lists and the 512 MiB "
                             "table: far beyond the 126 MB L2; no explicit flush",
                       "exchange": "none (1 GPU)" if world == 1 else (
                           ("fused NVLink peer kernel on byte planes: reduce-scatter + saturate + all-gather, in-kernel rank sync"
                            if os.environ.get("KMN_EXCHANGE_PLANES", "1") != "0" else
                            "fused NVLink peer kernel on u16 cells: reduce-scatter + saturate + all-gather, in-kernel rank sync") if fused
                           else "NCCL all-reduce of the u32 accumulators")},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_elapsed / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
        })
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true", help="resident leg only (for ncu): no e2e, no CPU baseline")
    args = ap.parse_args()
    claim_stdout()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        log("note: fewer than 3 warm-up steps requested; using 3")
        args.warmup = 3
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
