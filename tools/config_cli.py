#!/usr/bin/env python
"""A named BASELINE.json configuration through the product CLI: synthetic proteomes written as FASTA files,
`kamino-b200` on one GPU and on `--devices` sets, optionally the oracle CLI (C++ restatement of kamino 2.0.0)
beside them.  The four output files must be byte-identical across all runs; one JSON line with sha256, wall times
and the CLI's own stage times.

    python tools/config_cli.py --species 1000 --devices "0;0,1,2,3,4,5,6,7" --oracle     # configs[2]
    python tools/config_cli.py --species 1000 --nj --devices "0" --oracle                 # configs[4] shape
"""
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
    ap.add_argument("--species", type=int, default=1000)
    ap.add_argument("--kind", default="bacterial", choices=["bacterial", "eukaryotic"])
    ap.add_argument("--nj", action="store_true")
    ap.add_argument("--devices", default="0", help="device sets separated by ';', e.g. \"0;0,1,2,3\"")
    ap.add_argument("--oracle", action="store_true")
    ap.add_argument("--workdir", default="/tmp/kamino_config")
    ap.add_argument("--seed-offset", type=int, default=2)
    args = ap.parse_args()
    import oracle_ffi
    from kamino_b200 import build, synth
    cli = build.build_host()
    work = Path(args.workdir)
    shutil.rmtree(work, ignore_errors=True)
    work.mkdir(parents=True)
    t0 = time.time()
    species = getattr(synth, args.kind)(args.species, seed=synth.SEED0 + args.seed_offset)
    n_res = sum(int(s.lens.sum()) for s in species)
    synth.write_fasta_dir_fast(species, work / "in")
    del species
    out = {"what": "config_cli", "species": args.species, "kind": args.kind, "residues": n_res, "nj": args.nj,
           "generate_s": round(time.time() - t0, 1), "runs": []}
    threads = len(os.sched_getaffinity(0))
    common = ["-i", str(work / "in"), "-t", str(threads)] + (["--nj"] if args.nj else [])
    files = FILES if args.nj else FILES[:3]
    shas = []
    for devs in args.devices.split(";"):
        d = work / ("got_" + devs.replace(",", "_"))
        d.mkdir()
        env = dict(os.environ, KAMINO_TIMING="1")
        t0 = time.time()
        r = subprocess.run([str(cli), *common, "--devices", devs, "-o", str(d / "kamino")], capture_output=True, text=True, env=env)
        dt = time.time() - t0
        if r.returncode != 0:
            print(r.stderr, file=sys.stderr)
            raise SystemExit(f"kamino-b200 --devices {devs} failed")
        sha = {f: hashlib.sha256((d / f).read_bytes()).hexdigest()[:16] for f in files}
        shas.append(sha)
        aln = (d / FILES[0]).read_text().split(">")[1].splitlines()[1:]
        out["runs"].append({"impl": "kamino-b200", "devices": devs, "wall_s": round(dt, 2), "sha256": sha,
                            "alignment_length": len("".join(aln)),
                            "stage_times": [l[7:] for l in r.stderr.splitlines() if l.startswith("[time]")],
                            "summary": [l.strip() for l in r.stderr.splitlines() if "variant groups" in l or "alignment:" in l]})
    if args.oracle:
        oracle_ffi.build()
        d = work / "want"
        d.mkdir()
        t0 = time.time()
        r = subprocess.run([str(oracle_ffi.CLI), *common, "-o", str(d / "kamino")], capture_output=True, text=True)
        dt = time.time() - t0
        if r.returncode != 0:
            print(r.stderr, file=sys.stderr)
            raise SystemExit("oracle CLI failed")
        sha = {f: hashlib.sha256((d / f).read_bytes()).hexdigest()[:16] for f in files}
        shas.append(sha)
        out["runs"].append({"impl": "oracle CLI (C++ restatement of kamino 2.0.0)", "threads": threads, "wall_s": round(dt, 2), "sha256": sha})
    out["all_identical"] = all(s == shas[0] for s in shas)
    print(json.dumps(out))
    shutil.rmtree(work, ignore_errors=True)
    if not out["all_identical"]:
        raise SystemExit("outputs differ")


if __name__ == "__main__":
    main()
