#!/usr/bin/env python
"""bench.py — k-mer table build (pass 1) throughput on B200, BASELINE.json's metric.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference            # CPU arm: C++ restatement of kamino 2.0.0

A "step" is one pass of the hot path over one batch: 100 synthetic bacterial proteomes per GPU
(BASELINE.json configs[1]) -> recode/frame, exact per-species Bloom, count-min build into the
saturating u16 table.  The timed region is ONE JOB: the K steps accumulate into one table (kamino builds one
table per job, group_extraction.rs:527-660; from the second batch on the merge reads the table instead of
overwriting it), then the table is finalized — at N>1 by ONE table exchange (fused NVLink peer kernel, or NCCL
sum), inside the timed region.  `per_step_exchange` keeps round 1's definition (reset + exchange in every step).
  value     : residues/s, inputs already resident in HBM (kmn_count_staged + kmn_finalize_table)
  e2e       : residues/s through the C ABI from pinned HOST buffers (kmn_count_batch: H2D inside the
              timed region) + finalize + D2H of the per-species new-k-mer counts
  roofline  : whole pass 1 against the measured HBM copy bandwidth: algorithmic fraction (frac), hardware DRAM
              fraction (dram_fraction, from the committed ncu capture) and a per-kernel table
  pass2, nj : BASELINE.md 3's secondary metrics (occupancy scan residues/s, --nj pair-columns/s), N = 1
  checks    : asserted before timing: same table hash on every rank, row sums == new k-mers, and (N > 1) a
              small sharded case equal to the oracle's table
  strong_c2 : BASELINE.json configs[2] as ONE job (1,000 proteomes round-robin over the N GPUs, one exchange)
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
WORKLOAD = f"{N_SPECIES} synthetic bacterial proteomes (~4k proteins x ~330 aa), k={K}, {SCHEME_NAME}"   # identical in both arms
C2_SPECIES = int(os.environ.get("KMN_BENCH_C2_SPECIES", "1000"))   # BASELINE configs[2]: the strong-scaling job


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
        "config": {"workload": WORKLOAD,
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


def verify_table(ctx, world, rank, local_rank, new, exchange_and_finalize):
    """Asserted BEFORE anything is timed (the table is the one the last warm-up step left behind):
      * every rank holds the same table (64-bit position-sensitive hash, reduced on the device);
      * per count-min row, sum(cells) == sum over all ranks of the new k-mers (proba_filter.rs:146-168 adds 1 to one
        cell per row and new k-mer; holds while nothing saturates, which the bench workloads never do);
      * N > 1: a small species set sharded over the ranks and exchanged equals the single-process oracle table."""
    import torch
    import torch.distributed as dist
    rows, h = ctx.table_checksum()
    total_new = int(new.sum())
    hashes = [h]
    if world > 1:
        t = torch.tensor([total_new], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        total_new = int(t.item())
        hashes = [None] * world
        dist.all_gather_object(hashes, h)
    assert len(set(hashes)) == 1, f"ranks hold different tables after the exchange: {hashes}"
    assert [int(x) for x in rows] == [total_new] * 4, f"row sums {rows.tolist()} != new k-mers {total_new}"
    out = {"table_hash": f"{h:016x}", "identical_on_all_ranks": True, "row_sums_equal_new_kmers": total_new}
    if world > 1:
        sys.path.insert(0, str(ROOT / "tests"))
        import oracle_ffi
        from kamino_b200 import sharding, synth
        small = synth.stage(synth.bacterial(2 * world + 1, seed=99, n_prot=200, mean_len=150.0))
        mine = sharding.shard_species(small.n_sp, world, rank)
        res, rec, sp = sharding.select_species(small.residues, small.rec_off, small.sp_rec_off, mine)
        ctx.reset_table()
        ctx.count_batch(res, rec, sp)
        exchange_and_finalize()
        got = ctx.table_snapshot() if rank == 0 else None
        _, h_small = ctx.table_checksum()
        hs = [None] * world
        dist.all_gather_object(hs, h_small)
        assert len(set(hs)) == 1, "ranks hold different tables (small oracle case)"
        if rank == 0:
            orc = oracle_ffi.load()
            hh = orc.cmsa_new()
            try:
                orc.count_batch(hh, small.residues, small.rec_off, small.sp_rec_off, K, SCHEME, threads=4)
                assert np.array_equal(got, orc.cmsa_table(hh)), "exchanged table differs from the oracle's"
            finally:
                orc.cmsa_free(hh)
        # (the small batch stays staged: releasing would also drop the bench batch; it is a few hundred kB)
        out["small_case_equals_oracle"] = f"{small.n_sp} species over {world} ranks"
    log(f"[rank {rank}] table checks passed: {out}")
    return out


def secondary_legs(ctx, batch, h_res, h_rec, h_sp):
    """BASELINE.md 3 secondary metrics in the same JSON line: pass-2 residues/s (occupancy scan + boundary hits,
    then anchor pairing and grouping) on the bench workload and --nj pair-columns/s (n = 2000, L = 100000)."""
    sys.path.insert(0, str(ROOT / "tests"))
    ctx.reset_table()
    bid, _ = ctx.count_batch(h_res, h_rec, h_sp)
    ctx.finalize_table()
    min_needed = int(np.ceil(np.float32(0.85) * np.float32(batch.n_sp)))
    for _ in range(2):
        ns, ne = ctx.scan_batch(bid, min_needed)
    scan = []
    for _ in range(5):
        ctx.scan_batch(bid, min_needed)
        scan.append(ctx.stage_ms("scan"))
    pair = []
    for _ in range(3):
        n_pairs = ctx.pair_batch(bid, 35)
        pair.append(ctx.stage_ms("pair_batch"))
    grp_ms, n_groups, n_valid = None, None, None
    if hasattr(ctx, "group_batches"):
        for _ in range(3):
            n_obs, n_groups, n_valid = ctx.group_batches([bid], min_needed)
            grp_ms = ctx.stage_ms("group_observations")
    scan_ms, pair_ms = float(np.median(scan)), float(np.median(pair))
    pass2 = {"value": batch.n_residues / (scan_ms * 1e-3), "unit": UNIT, "scan_ms": scan_ms, "starts": int(ns), "ends": int(ne),
             "pair_batch_ms": pair_ms, "anchor_pairs": int(n_pairs), "group_ms": grp_ms, "groups": n_groups,
             "groups_passing_length_selection": n_valid, "min_needed": min_needed,
             "algorithmic_bytes_per_residue": 1 + 8 * 0.96 + 2 * 0.96}
    ctx.release_batches()
    rng = np.random.default_rng(1)
    n, L = 2000, 100_000
    from kamino_b200 import synth
    base = rng.choice(20, size=L, p=synth.LG_PI).astype(np.uint8)
    m = np.tile(base, (n, 1))
    mut = rng.random((n, L)) < 0.05
    m[mut] = rng.integers(0, 20, int(mut.sum()), dtype=np.uint8)
    m[rng.random((n, L)) < 0.05] = 20
    ctx.pair_counts(m[:64])
    ks, cs = [], []
    for _ in range(3):
        t0 = time.perf_counter()
        ctx.pair_counts(m)
        cs.append((time.perf_counter() - t0) * 1e3)
        ks.append(ctx.stage_ms("pair_counts"))
    k_ms, c_ms = float(np.median(ks)), float(np.median(cs))
    pairs = n * (n - 1) // 2
    nj = {"value": pairs * L / (k_ms * 1e-3), "unit": "pair-columns/s", "n": n, "L": L, "kernel_ms": k_ms, "call_ms": c_ms,
          "call_value": pairs * L / (c_ms * 1e-3)}
    return pass2, nj


def strong_scaling_c2(ctx, world, rank, local_rank, barrier, exchange_and_finalize, args):
    """BASELINE.json configs[2]: ONE job of 1,000 synthetic bacterial proteomes, species round-robin over the
    ranks, one table exchange per job.  Reported next to the weak-scaling headline (driver: strong_c2.value at
    N vs N = 1).  Every rank generates the same species set (same seed) and keeps its share."""
    import torch
    import torch.distributed as dist
    from kamino_b200 import sharding, synth
    t0 = time.time()
    species = synth.bacterial(C2_SPECIES, seed=synth.SEED0 + 2)
    mine = [species[i] for i in sharding.shard_species(len(species), world, rank)]
    n_res_job = int(sum(int(s.lens.sum()) for s in species))
    del species
    batch = synth.stage(mine)
    log(f"[rank {rank}] strong C2: {len(mine)} of {C2_SPECIES} species, {batch.n_residues} residues ({time.time() - t0:.0f}s to generate)")
    pin_res = torch.empty(batch.residues.size, dtype=torch.uint8).pin_memory()
    pin_res.numpy()[:] = batch.residues
    h_res = pin_res.numpy()
    pin_rec = torch.from_numpy(batch.rec_off.astype(np.int64)).pin_memory()
    pin_sp = torch.from_numpy(batch.sp_rec_off.astype(np.int64)).pin_memory()
    h_rec, h_sp = pin_rec.numpy().view(np.uint64), pin_sp.numpy().view(np.uint64)
    steps = max(1, min(args.steps, 5))
    lib_stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local_rank))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, n):
        barrier()
        ev0.record(lib_stream)
        for _ in range(n):
            fn()
        ev1.record(lib_stream)
        barrier()
        dt = ev0.elapsed_time(ev1) * 1e-3
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt

    bid = ctx.stage_batch(h_res, h_rec, h_sp)
    state = {}

    def job_resident():
        ctx.reset_table()
        state["new"] = ctx.count_staged(bid, batch.n_sp)
        exchange_and_finalize()

    def job_e2e():
        ctx.reset_table()
        _, state["new"] = ctx.count_batch(h_res, h_rec, h_sp)
        exchange_and_finalize()
        ctx.release_batches()

    job_resident()
    # the job's table: identical on every rank, row sums == new k-mers of the whole job
    rows, h = ctx.table_checksum()
    tot = torch.tensor([int(state["new"].sum())], device="cuda", dtype=torch.int64)
    hs = [h]
    if world > 1:
        dist.all_reduce(tot)
        hs = [None] * world
        dist.all_gather_object(hs, h)
    assert len(set(hs)) == 1, "strong C2: ranks hold different tables"
    assert [int(x) for x in rows] == [int(tot.item())] * 4, "strong C2: row sums != new k-mers"
    dt = timed(job_resident, steps)
    ctx.release_batches()
    job_e2e()
    rows_e, h_e = ctx.table_checksum()   # the job from host buffers (long range schedule) builds the same table
    assert h_e == h and [int(x) for x in rows_e] == [int(x) for x in rows], "strong C2: the e2e job's table differs from the resident job's"
    dt_e2e = timed(job_e2e, steps)
    return {"workload": f"{C2_SPECIES} synthetic bacterial proteomes (~4k proteins x ~330 aa), k={K}, {SCHEME_NAME}, "
                        f"species round-robin over {world} GPU(s), one table exchange per job",
            "scaling": "strong", "residues": n_res_job, "jobs_timed": steps,
            "value": n_res_job * steps / dt, "unit": UNIT, "ms_per_job": 1e3 * dt / steps,
            "e2e": {"value": n_res_job * steps / dt_e2e, "unit": UNIT, "ms_per_job": 1e3 * dt_e2e / steps,
                    "h2d_bytes_per_job_per_gpu": int(batch.residues.size + h_rec.nbytes + h_sp.nbytes)},
            "table_hash": f"{h:016x}", "row_sums_equal_new_kmers": int(tot.item())}


def bind_to_gpu_numa_node(local_rank):
    """N > 1: run this rank (and so first-touch its pinned buffers) on the CPUs NVML reports as local to its GPU.
    Without it every rank's pinned memory may sit on one socket and the per-GPU host-to-device rate of the e2e leg
    drops from 55 to ~20 GB/s at N = 8 (round 1's SCALE record).  Returns the CPU list, None when nothing was done."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        p = torch.cuda.get_device_properties(local_rank)
        try:
            h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{p.pci_domain_id:08x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0")
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = (max(os.cpu_count() or 1, 1) + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception as e:   # best effort: NVML missing, no permission, a container without the topology
        log(f"[bench] no NUMA binding ({type(e).__name__}: {e})")
        return None


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
    numa_cpus = None
    if world > 1 and os.environ.get("KMN_BENCH_NUMA", "1") != "0":
        numa_cpus = bind_to_gpu_numa_node(local_rank)
        log(f"[rank {rank}] bound to the {len(numa_cpus)} CPUs local to GPU {local_rank}" if numa_cpus else f"[rank {rank}] not bound to a NUMA node")
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

    def step_resident():                                 # one batch into the job's table
        return ctx.count_staged(bid, batch.n_sp)

    def step_e2e():
        _, new = ctx.count_batch(h_res, h_rec, h_sp)     # H2D of residues + offsets inside
        return new                                       # D2H: per-species counts

    def job(step, n, after_step=None):                   # n batches into one table, one finalize / exchange
        ctx.reset_table()
        new = None
        for _ in range(n):
            new = step()
            if after_step:
                after_step()
        exchange_and_finalize()
        return new

    def isolated_step():                                 # round 1's step: a job of one batch
        return job(step_resident, 1)

    sampler = ClockSampler(local_rank)
    sampler.start()                      # comes up during the warm-up; only samples inside the timed regions count
    new = job(step_resident, max(args.warmup, 1))
    new = isolated_step()                # leaves the table of exactly one batch for the checks
    u = float(new.sum()) / batch.n_residues
    checks = verify_table(ctx, world, rank, local_rank, new, exchange_and_finalize)
    sampler.wait_first()

    # ---- timed: resident inputs
    stage_acc = {"frame": 0.0, "bloom": 0.0, "cms": 0.0, "finalize": 0.0, "cms_partition_kernel": 0.0,
                 "bloom_bucket_kernel": 0.0, "bloom_partition_kernel": 0.0, "cms_apply_kernel": 0.0}

    def add_stage_times():
        for st in stage_acc:
            stage_acc[st] += ctx.stage_ms(st)

    launches0 = ctx.launch_count()
    lib_stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local_rank))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed_region(fn):
        barrier()
        sampler.begin()
        ev0.record(lib_stream)
        fn()
        ev1.record(lib_stream)
        barrier()
        sampler.end()
        dt = max(ev0.elapsed_time(ev1) * 1e-3, 1e-9)   # device clock, CUDA events on the launching stream
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt

    elapsed = timed_region(lambda: job(step_resident, args.steps, add_stage_times))
    launches = ctx.launch_count() - launches0
    value = n_res_total * args.steps / elapsed
    per_step_exchange = None
    if world > 1 and not args.profile:   # round 1's definition, for continuity: every step resets and exchanges
        def isolated_steps():
            for _ in range(args.steps):
                isolated_step()
        dt = timed_region(isolated_steps)
        per_step_exchange = {"value": n_res_total * args.steps / dt, "unit": UNIT, "ms_per_step": 1e3 * dt / args.steps,
                             "note": "every step = reset + 100 proteomes per GPU + table exchange (a job of one batch)"}

    # ---- timed: end to end from pinned host buffers
    ctx.release_batches()
    e2e_steps = 0 if args.profile else args.steps
    if e2e_steps:
        job(step_e2e, 1, ctx.release_batches)
    link_gbs = []

    def after_e2e_step():                                # host side only: the rate the library measured over the step's copies
        link_gbs.append(ctx.stage_ms("h2d_gbs"))
        ctx.release_batches()

    # (a) one synchronous kmn_count_batch per batch: copy / compute ranges overlap INSIDE a call, nothing across calls
    e2e_sync_elapsed = timed_region(lambda: job(step_e2e, e2e_steps, after_e2e_step)) if e2e_steps else 1e-9

    # (b) the multi-batch job as the C ABI means it to be driven: kmn_stage_batch_async(batch i+1) is queued before
    # kmn_count_staged(batch i), so the upload of the next batch runs under the kernels of the current one.  Every batch
    # is copied from pinned host memory inside the timed region (two alternating device batches), every step reads its
    # per-species counts back.
    def job_e2e_pipelined(n):
        ctx.reset_table()
        new = None
        nxt = ctx.stage_batch_async(h_res, h_rec, h_sp)
        for i in range(n):
            cur = nxt
            if i + 1 < n:
                nxt = ctx.stage_batch_async(h_res, h_rec, h_sp)
            new = ctx.count_staged(cur, batch.n_sp)
            ctx.release_batch(cur)
        exchange_and_finalize()
        ctx.release_batches()
        return new

    if e2e_steps:
        new_p = job_e2e_pipelined(2)
        assert new_p.tolist() == new.tolist(), "pipelined e2e job: per-species counts differ from the resident step's"
    e2e_elapsed = timed_region(lambda: job_e2e_pipelined(e2e_steps)) if e2e_steps else 1e-9
    clocks = sampler.stop()
    e2e_value = n_res_total * args.steps / e2e_elapsed
    h2d = int(batch.residues.size + batch.rec_off.nbytes + batch.sp_rec_off.nbytes)
    d2h = int(batch.n_sp * 8 + (batch.n_sp + 1) * 4)

    # ---- roofline (BASELINE.md 3 / SURVEY.md 8d): the WHOLE pass over the measured copy bandwidth.
    # Algorithmic bytes of pass 1: B1 = 1 + 16u per residue (1 residue byte + 4 cells x (2 B read + 2 B write) per
    # k-mer that passes the Bloom filter; u measured).  frac = B1 * residues / step time / peak.  dram_fraction is
    # the same with the bytes the DRAM actually moved (sum of ncu dram__bytes_{read,write}.sum over the step's
    # kernels, captured once per code change: profiles/roofline_traffic.json).  The per-kernel table charges every
    # kernel ITS OWN algorithmic bytes against ITS OWN time (CUDA events of the library around single launches).
    peak, peak_src = measured_peaks()
    stage_ms = {s: v / args.steps for s, v in stage_acc.items()}
    step_ms = 1e3 * elapsed / args.steps
    n_res, n_pos = batch.n_residues, batch.n_residues + len(batch.rec_off) - 1
    alg_bytes = (1.0 + 16.0 * u) * n_res
    traffic_tab, traffic_src = {}, None
    tp = ROOT / "profiles" / "roofline_traffic.json"
    if tp.exists():
        try:
            traffic_tab = json.loads(tp.read_text())
            traffic_src = traffic_tab.get("_source")
        except Exception:
            traffic_tab = {}
    own = {   # algorithmic bytes of one launch of each kernel group (own inputs + own outputs, no sector rounding)
        "frame": ("frame_kernel (+ tile records, codes tail)", float(batch.residues.size) + n_pos,        # raw bytes read once, codes written
                  ["frame_kernel", "frame_tile_records_kernel", "codes_tail_kernel"]),
        "bloom_partition_kernel": ("bloom_partition_kernel", n_pos * (1 + 8.0), ["bloom_partition_kernel"]),          # code in, 2 probes x 4 B out
        "bloom_bucket_kernel": ("bloom_bucket_kernel", n_pos * (8.0 + 0.25), ["bloom_bucket_kernel"]),                # probes in, 2 lost bits out
        "cms_partition_kernel": ("cms_partition_kernel", n_pos * 1.25 + 8.0 * u * n_res, ["cms_partition_kernel"]),   # code + lost bits in, 4 x 2 B entries out
        "cms_apply_kernel": ("cms_apply_kernel", 8.0 * u * n_res + 4.0 * ffi.CMS_CELLS, ["cms_apply_kernel"]),        # entries in, every touched cell read and written
    }
    kernels = []
    for key, (name, nbytes, parts) in own.items():
        ms = stage_ms[key]
        tr = None
        for part in parts:
            if part in traffic_tab:
                tr = (tr or 0) + traffic_tab[part]["dram_bytes_per_launch"]
        kernels.append({"kernel": name, "ms": ms, "share_of_step": ms / step_ms if step_ms else None,
                        "algorithmic_bytes": nbytes, "achieved": nbytes / (ms * 1e-3) / 1e9 if ms else None,
                        "frac": nbytes / (ms * 1e-3) / 1e9 / peak if ms else None, "traffic": tr,
                        "dram_fraction": tr / (ms * 1e-3) / 1e9 / peak if (tr and ms) else None})
    traffic = sum(k["traffic"] for k in kernels if k["traffic"]) if any(k["traffic"] for k in kernels) else None
    dominant = max(kernels, key=lambda k: k["ms"])
    achieved = alg_bytes / (step_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "kernel": "pass 1: every kernel of a step (frame, Bloom partition + buckets, count-min partition + apply)",
        "launches_per_step": int(launches // max(args.steps, 1)),
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
        "dram_fraction": traffic / (step_ms * 1e-3) / 1e9 / peak if traffic else None,
        "traffic_over_algorithmic": traffic / alg_bytes if traffic else None, "traffic_source": traffic_src,
        "peak_source": peak_src, "algorithmic_bytes_per_residue": 1.0 + 16.0 * u, "algorithmic_bytes_per_step": alg_bytes,
        "new_kmer_fraction_u": u, "ms_per_step": step_ms, "dominant_kernel": dominant["kernel"], "kernels": kernels,
        "limiter": "instruction issue and shared-memory atomics, not DRAM bytes (ncu --set full summaries under profiles/): "
                   "the roofline is stated against HBM because the path has no floating-point work",
    }

    # ---- secondary legs of BASELINE.md 3 (same process, after the timed regions): pass 2 and --nj
    pass2 = nj = None
    if world == 1 and not args.profile:
        pass2, nj = secondary_legs(ctx, batch, h_res, h_rec, h_sp)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.profile:
        threads = host_threads()
        n_res_c, dt, _, n_sp = cpu_pass1(batch, threads, n_species=8)
        n_sample = int(max(8, min(N_SPECIES, 15.0 / max(dt / n_sp, 1e-6))))
        n_res_c, dt, _, n_sp = cpu_pass1(batch, threads, n_species=n_sample)
        cpu_baseline = {"value": n_res_c / dt, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"{n_sp} of {N_SPECIES} species ({n_res_c} residues), pass 1 + snapshot, "
                                  "C++ restatement of kamino 2.0.0"}

    strong = None
    if args.strong:
        ctx.release_batches()
        strong = strong_scaling_c2(ctx, world, rank, local_rank, barrier, exchange_and_finalize, args)

    if rank == 0:
        emit({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "per_gpu": True,
                       "job": f"the {args.steps} timed steps are one job: every step adds its batch to the job's table, the table is "
                              "finalized once (N > 1: one table exchange), all inside the timed region",
                       "residues_per_gpu": batch.n_residues, "species_per_gpu": batch.n_sp,
                       "l2": "per step the kernels stream 121 MB of residues, ~5.5 GB of partition runs / slots and the 512 MiB "
                             "table: far beyond the 126 MB L2; no explicit flush",
                       "exchange": "none (1 GPU)" if world == 1 else (
                           ("fused NVLink peer kernel on byte planes: reduce-scatter + saturate + all-gather, in-kernel rank sync"
                            if os.environ.get("KMN_EXCHANGE_PLANES", "1") != "0" else
                            "fused NVLink peer kernel on u16 cells: reduce-scatter + saturate + all-gather, in-kernel rank sync") if fused
                           else "NCCL all-reduce of the u32 accumulators")},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_elapsed / args.steps,
                    "how": "kmn_stage_batch_async(batch i+1) queued before kmn_count_staged(batch i): the upload of the next "
                           "batch overlaps the kernels of the current one; every batch copied from pinned host memory and "
                           "every step's counts read back inside the timed region",
                    "sync_call": {"value": n_res_total * args.steps / e2e_sync_elapsed, "unit": UNIT,
                                  "ms_per_step": 1e3 * e2e_sync_elapsed / args.steps,
                                  "how": "one synchronous kmn_count_batch per batch (copy / compute ranges overlap inside a "
                                         "call, nothing across calls): what a job of ONE batch pays"},
                    "h2d_gbs": {"median": float(np.median(link_gbs)), "min": float(np.min(link_gbs)),
                                "note": "host-to-device rate of rank 0's range copies as the library measured it per step "
                                        "(its range schedule adapts to it; below 40 GB/s it switches to equal ranges)"}
                    if link_gbs else None},
            "gpu_launches": int(launches),
            "host_binding": (f"rank 0 on the {len(numa_cpus)} CPUs NVML reports local to its GPU (every rank likewise)" if numa_cpus
                             else "none"),
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "pass2": pass2,
            "nj": nj,
            "checks": checks,
            "per_step_exchange": per_step_exchange,
            "strong_c2": strong,
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
    ap.add_argument("--strong", dest="strong", action="store_true", default=None,
                    help="also run the strong-scaling job of BASELINE configs[2] (1,000 proteomes over the N GPUs); "
                         "default: on (KMN_BENCH_STRONG=0 or --no-strong turns it off)")
    ap.add_argument("--no-strong", dest="strong", action="store_false")
    args = ap.parse_args()
    claim_stdout()
    if args.strong is None:
        args.strong = os.environ.get("KMN_BENCH_STRONG", "1") != "0"
    if args.profile:
        args.strong = False
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        log("note: fewer than 3 warm-up steps requested; using 3")
        args.warmup = 3
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
