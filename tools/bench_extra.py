#!/usr/bin/env python
"""Secondary measurements: pass-2 occupancy scan and the --nj pair-count kernel, each
next to the CPU oracle on a bounded sample.  Writes one JSON object per line to stdout."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--species", type=int, default=100)
    ap.add_argument("--fasta", action="store_true", help="also time kmn_count_fasta on the FASTA text of the species")
    ap.add_argument("--nj-n", type=int, default=2000)
    ap.add_argument("--nj-l", type=int, default=100_000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--euk", type=int, default=0, help="also run pass 1 on N synthetic eukaryotic proteomes "
                    "(BASELINE.json configs[3] shape, ~10 M residues each) and compare the table with the oracle")
    args = ap.parse_args()
    import oracle_ffi
    from kamino_b200 import ffi, synth
    orc = oracle_ffi.load()
    threads = len(os.sched_getaffinity(0))

    # ---- pass 2 on the bacterial config
    species = synth.bacterial(args.species)
    b = synth.stage(species)
    min_needed = orc.min_needed(0.85, b.n_sp)
    with ffi.Context(13, "sr6") as ctx:
        bid, new = ctx.count_batch(b.residues, b.rec_off, b.sp_rec_off)
        ctx.finalize_table()
        table = ctx.table_snapshot()
        for _ in range(2):
            ns, ne = ctx.scan_batch(bid, min_needed)
        ms = []
        for _ in range(args.reps):
            ctx.scan_batch(bid, min_needed)
            ms.append(ctx.stage_ms("scan"))
        scan_ms = float(np.median(ms))
        pm = []
        for _ in range(args.reps):
            n_pairs = ctx.pair_batch(bid, 35)
            pm.append(ctx.stage_ms("pair_batch"))
        pair_ms = float(np.median(pm))
        # grouping of all anchor pairs by (left, right) + per-group length selection on the device
        allp = [ctx.fetch_pairs(bid, s) for s in range(b.n_sp)]
        sid_all = np.concatenate([np.full(len(p), s, np.uint32) for s, p in enumerate(allp)])
        allp = np.concatenate(allp)
        gm = []
        for _ in range(args.reps):
            perm, groups = ctx.group_observations(allp["left"], allp["right"], sid_all, allp["len"], min_needed)
            gm.append(ctx.stage_ms("group_observations"))
        group_ms = float(np.median(gm))
        n_valid = int(np.count_nonzero(groups["flags"] & 1))
        # CPU: oracle scan of a few species (single thread per species, run sequentially)
        cms = orc.cms_from_table(table)
        n_cpu = min(4, b.n_sp)
        t0 = time.perf_counter()
        res_cpu = 0
        for s in range(n_cpu):
            ws, we, _ = orc.scan_species(cms, b.residues, b.rec_off, int(b.sp_rec_off[s]), int(b.sp_rec_off[s + 1]),
                                         13, 1, min_needed, want_occ=False)
            res_cpu += int(species[s].lens.sum())
            gs, ge = ctx.fetch_hits(bid, s)
            assert np.array_equal(gs, ws) and np.array_equal(ge, we)
            wp = orc.pair_species(cms, b.residues, b.rec_off, int(b.sp_rec_off[s]), int(b.sp_rec_off[s + 1]), 13, 1,
                                  min_needed, 35)
            assert np.array_equal(ctx.fetch_pairs(bid, s), wp)
        cpu_dt = time.perf_counter() - t0
        orc.cms_free(cms)
        print(json.dumps({"what": "pass2_scan", "species": b.n_sp, "residues": b.n_residues, "ms": scan_ms,
                          "residues_per_s": b.n_residues / (scan_ms * 1e-3), "starts": ns, "ends": ne,
                          "pair_batch_ms": pair_ms, "anchor_pairs": n_pairs,
                          "group_kernels_ms": group_ms, "groups": int(len(groups)), "groups_passing_length_selection": n_valid,
                          "cpu_single_thread_residues_per_s": res_cpu / cpu_dt, "parity_species_checked": n_cpu}))

    # ---- FASTA text in, table out: record detection on the device (kmn_count_fasta) next to the host-parsed batch
    if args.fasta:
        import tempfile
        parts, file_off = [], [0]
        for sid, sp_ in enumerate(species):
            chunks, pos = [], 0
            for pid, ln in enumerate(sp_.lens):
                chunks.append(f">sp{sid}_p{pid} hypothetical protein {pid} [Synth sp]\n".encode())
                seq = sp_.seq[pos:pos + ln].tobytes()
                pos += int(ln)
                chunks += [seq[i:i + 60] + b"\n" for i in range(0, len(seq), 60)]
            parts.append(b"".join(chunks))
            file_off.append(file_off[-1] + len(parts[-1]))
        file_off = np.asarray(file_off, np.uint64)
        text = ffi.PinnedBuffer(int(file_off[-1]))
        text.array[:] = np.frombuffer(b"".join(parts), np.uint8)
        pin = ffi.PinnedBuffer(b.residues.size)
        pin.array[:] = b.residues
        with tempfile.NamedTemporaryFile(suffix=".faa") as f:   # the CPU parser on one proteome
            f.write(parts[0])
            f.flush()
            t0 = time.perf_counter()
            orc.read_fasta(f.name)
            parse_dt = time.perf_counter() - t0
        with ffi.Context(13, "sr6") as ctx:
            t_fa, t_batch = [], []
            for rep in range(args.reps + 1):
                ctx.reset_table(); ctx.synchronize()
                t0 = time.perf_counter()
                _, new_fa, n_rec = ctx.count_fasta(text.array, file_off)
                t_fa.append((time.perf_counter() - t0) * 1e3)
                ctx.finalize_table()
                tab_fa = ctx.table_snapshot() if rep == 0 else None
                ctx.release_batches()
                ctx.reset_table(); ctx.synchronize()
                t0 = time.perf_counter()
                _, new_b = ctx.count_batch(pin.array, b.rec_off, b.sp_rec_off)
                t_batch.append((time.perf_counter() - t0) * 1e3)
                ctx.finalize_table()
                if rep == 0:
                    assert np.array_equal(ctx.table_snapshot(), tab_fa) and new_fa.tolist() == new_b.tolist()
                ctx.release_batches()
        fa_ms, batch_ms = float(np.median(t_fa[1:])), float(np.median(t_batch[1:]))
        print(json.dumps({"what": "count_fasta", "species": b.n_sp, "fasta_bytes": int(file_off[-1]), "records": n_rec,
                          "residues": b.n_residues, "count_fasta_ms": fa_ms, "residues_per_s": b.n_residues / (fa_ms * 1e-3),
                          "count_batch_ms_host_parsed_input": batch_ms,
                          "cpu_parse_one_proteome_ms_single_thread": parse_dt * 1e3, "table_equal": True}))
        text.free(); pin.free()

    # ---- pass 1 on the eukaryotic shape (long proteomes: ~1200 Bloom tiles per species, visible Bloom FP rate)
    if args.euk:
        species = synth.eukaryotic(args.euk)
        b = synth.stage(species)
        with ffi.Context(13, "sr6") as ctx:
            bid = ctx.stage_batch(b.residues, b.rec_off, b.sp_rec_off)
            for _ in range(2):
                ctx.reset_table()
                new = ctx.count_staged(bid, b.n_sp)
            ms = sum(ctx.stage_ms(s) for s in ("frame", "bloom", "cms"))
            ctx.finalize_table()
            table = ctx.table_snapshot()
        h = orc.cmsa_new()
        t0 = time.perf_counter()
        want_new = orc.count_batch(h, b.residues, b.rec_off, b.sp_rec_off, 13, 1, threads=threads)
        cpu_dt = time.perf_counter() - t0
        same = bool(np.array_equal(table, orc.cmsa_table(h))) and new.tolist() == want_new.tolist()
        orc.cmsa_free(h)
        print(json.dumps({"what": "pass1_eukaryotic", "species": b.n_sp, "residues": b.n_residues, "ms": ms,
                          "residues_per_s": b.n_residues / (ms * 1e-3), "passing_fraction": float(new.sum()) / b.n_residues,
                          "cpu_residues_per_s": b.n_residues / cpu_dt, "cpu_threads": threads, "table_equals_oracle": same}))
        assert same

    # ---- --nj pair counts
    rng = np.random.default_rng(1)
    n, L = args.nj_n, args.nj_l
    base = rng.choice(20, size=L, p=synth.LG_PI).astype(np.uint8)
    m = np.tile(base, (n, 1))
    mut = rng.random((n, L)) < 0.05
    m[mut] = rng.integers(0, 20, int(mut.sum()), dtype=np.uint8)
    m[rng.random((n, L)) < 0.05] = 20
    with ffi.Context(13, "sr6") as ctx:
        ctx.pair_counts(m[:64])
        ms = []
        for _ in range(args.reps):
            v, mm = ctx.pair_counts(m)
            ms.append(ctx.stage_ms("pair_counts"))
        k_ms = float(np.median(ms))
    n_cpu = 48
    t0 = time.perf_counter()
    wv, wm = orc.pair_counts(m[:n_cpu], threads=threads)
    cpu_dt = time.perf_counter() - t0
    iu = np.triu_indices(n_cpu, 1)
    assert np.array_equal(v[:n_cpu, :n_cpu][iu], wv[iu]) and np.array_equal(mm[:n_cpu, :n_cpu][iu], wm[iu])
    pairs = n * (n - 1) // 2
    print(json.dumps({"what": "pair_counts", "n": n, "L": L, "kernel_ms": k_ms,
                      "pair_columns_per_s": pairs * L / (k_ms * 1e-3),
                      "cpu_pair_columns_per_s": (n_cpu * (n_cpu - 1) // 2) * L / cpu_dt, "cpu_threads": threads}))


if __name__ == "__main__":
    main()
