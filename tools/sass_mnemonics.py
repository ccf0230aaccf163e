#!/usr/bin/env python
"""profiles/sass/: per-kernel counts of the Blackwell/Hopper-path SASS mnemonics of the built library and the SASS of
the hot kernels.   python tools/sass_mnemonics.py [lib]"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
lib = sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "kamino_b200" / "_lib" / "libkamino_b200.so")
WATCH = ["UTMASTG", "UTMALDG", "UBLKCP", "UTMACMDFLUSH", "SYNCS", "LDGSTS", "ATOMS", "ATOMG", "REDG", "RED", "REDUX", "BAR", "CCTL", "UTCMMA", "LDTM"]
HOT = ["frame_kernel", "bloom_partition_kernel", "bloom_bucket_kernel", "cms_partition_kernel", "cms_apply_kernel",
       "occupancy_sweep8_kernel"]

sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
funcs = OrderedDict()
cur, it = None, iter(names)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = re.sub(r"\((unsigned int|int|bool)\)", "", next(it)).split("(")[0]
        funcs[cur] = []
    elif cur is not None:
        funcs[cur].append(line)
out = ROOT / "profiles" / "sass"
out.mkdir(parents=True, exist_ok=True)
lines = ["# cuobjdump -sass kamino_b200/_lib/libkamino_b200.so: instruction count and Blackwell/Hopper-path mnemonics per kernel",
         "# (UTMASTG = TMA tensor store, UBLKCP = bulk async copy, SYNCS = mbarrier, LDGSTS = cp.async, ATOMS = shared-memory atomic)"]
for name, body in funcs.items():
    ops = [m.group(1) for l in body for m in [re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)] if m]
    c = Counter(o.split(".")[0] for o in ops)
    seen = "  ".join(f"{w}={c[w]}" for w in WATCH if c[w])
    lines.append(f"{name:<58} total={len(ops):5d}  {seen}")
    for h in HOT:
        if re.search(rf"\b{h}\b", name):
            tag = re.sub(r"[^0-9A-Za-z]+", "_", name.split("kmn::")[-1]).strip("_")
            (out / f"r02_{tag}.sass").write_text(f"// {name}\n" + "\n".join(l for l in body if not re.match(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", l)) + "\n")
(out / "r02_mnemonics.txt").write_text("\n".join(lines) + "\n")
print("\n".join(l for l in lines if any(h in l for h in HOT) or l.startswith("#")))
