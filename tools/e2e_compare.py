#!/usr/bin/env python
"""Full-pipeline comparison on synthetic proteomes written as FASTA files: `kamino-b200` (GPU front
end) vs the oracle CLI (C++ restatement of kamino 2.0.0, -t = host cores).  Outputs must be
byte-identical; wall times are end to end (gunzip + parse + both passes + grouping + writers)."""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
FILES = ["kamino_alignment.fas", "kamino_missing.tsv", "kamino_partitions.tsv", "kamino_NJ.tree"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--species", type=int, default=100)
    ap.add_argument("--kind", default="bacterial", choices=["bacterial", "eukaryotic"])
    ap.add_argument("--nj", action="store_true")
    ap.add_argument("--workdir", default="/tmp/kamino_e2e")
    ap.add_argument("--skip-oracle", action="store_true")
    ap.add_argument("--devices", default="", help="comma-separated CUDA ordinals for kamino-b200 --devices")
    args = ap.parse_args()
    import oracle_ffi
    from kamino_b200 import build, synth
    oracle_ffi.build()
    cli = build.build_host()
    work = Path(args.workdir)
    shutil.rmtree(work, ignore_errors=True)
    (work / "got").mkdir(parents=True)
    (work / "want").mkdir(parents=True)
    t0 = time.time()
    species = getattr(synth, args.kind)(args.species)
    synth.write_fasta_dir(species, work / "in")
    n_res = sum(int(s.lens.sum()) for s in species)
    del species
    gen_s = time.time() - t0
    threads = len(os.sched_getaffinity(0))
    common = ["-i", str(work / "in"), "-t", str(threads)] + (["--nj"] if args.nj else [])
    t0 = time.time()
    dev = ["--devices", args.devices] if args.devices else []
    env = dict(os.environ, KAMINO_TIMING="1")
    r = subprocess.run([str(cli), *common, *dev, "-o", str(work / "got" / "kamino")], capture_output=True, text=True, env=env)
    gpu_s = time.time() - t0
    if r.returncode != 0:
        print(r.stderr, file=sys.stderr)
        raise SystemExit("kamino-b200 failed")
    out = {"what": "e2e_cli", "species": args.species, "kind": args.kind, "residues": n_res, "threads": threads,
           "generate_s": round(gen_s, 1), "gpu_cli_s": round(gpu_s, 2), "devices": args.devices or "0",
           "gpu_stage_times": [l for l in r.stderr.splitlines() if l.startswith("[time]")]}
    if not args.skip_oracle:
        t0 = time.time()
        r = subprocess.run([str(oracle_ffi.CLI), *common, "-o", str(work / "want" / "kamino")], capture_output=True, text=True)
        out["oracle_cli_s"] = round(time.time() - t0, 2)
        if r.returncode != 0:
            print(r.stderr, file=sys.stderr)
            raise SystemExit("oracle failed")
        files = FILES if args.nj else FILES[:3]
        same = {f: (work / "got" / f).read_bytes() == (work / "want" / f).read_bytes() for f in files}
        out["identical"] = same
        out["all_identical"] = all(same.values())
    out["sha256"] = {f: hashlib.sha256((work / "got" / f).read_bytes()).hexdigest()[:16]
                     for f in FILES if (work / "got" / f).exists()}
    out["alignment_length"] = len("".join((work / "got" / FILES[0]).read_text().split(">")[1].splitlines()[1:]))
    print(json.dumps(out))
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
