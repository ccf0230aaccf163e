#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into one line per captured launch."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "dram_rd"),
        ("dram__bytes_write.sum", "dram_wr"), ("lts__t_sector_hit_rate.pct", "l2hit%"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__inst_executed.sum", "inst"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("launch__registers_per_thread", "regs"),
        ("l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum", "atom"), ("l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum", "red")]
idx = [(hdr.index(k), n) for k, n in want if k in hdr]
for r in rows[2:]:
    parts = []
    for i, n in idx:
        v = r[i]
        if n == "kernel":
            v = v.split("(")[0]
        parts.append(f"{n}={v}{'' if n in ('kernel',) else units[i] if units[i] not in ('%',) else ''}")
    print("  ".join(parts))
