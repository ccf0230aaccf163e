#!/usr/bin/env python
"""profiles/roofline_traffic.json (what bench.py's `traffic` / `dram_fraction` read) from ncu --set full reports:
per kernel dram__bytes_read.sum + dram__bytes_write.sum and the duration of ONE launch.  Later reports override
earlier ones (e.g. a second capture of cms_apply_kernel merging into a live table).

    python tools/traffic_from_ncu.py "<what was captured>" a.ncu-rep [b.ncu-rep ...]
"""
import csv, json, re, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def to_us(v, unit):
    f = float(v.replace(",", ""))
    return f * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}[unit]


def main():
    source, reps = sys.argv[1], sys.argv[2:]
    tab = {}
    for rep in reps:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        col = {k: hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum")}
        for r in rows[2:]:
            name = re.sub(r"^void\s+", "", r[col["Kernel Name"]].split("(")[0]).split("<")[0].strip()
            rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
            wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
            tab[name] = {"dram_bytes_per_launch": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr),
                         "ncu_duration_us": round(to_us(r[col["gpu__time_duration.sum"]], units[col["gpu__time_duration.sum"]]), 3),
                         "report": Path(rep).name}
    tab["_source"] = source
    tab["_sum_pass1_bytes"] = sum(v["dram_bytes_per_launch"] for k, v in tab.items() if not k.startswith("_"))
    (ROOT / "profiles" / "roofline_traffic.json").write_text(json.dumps(tab, indent=1) + "\n")
    print(json.dumps(tab, indent=1))


if __name__ == "__main__":
    main()
