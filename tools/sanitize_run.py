"""Small run of every kernel for compute-sanitizer (memcheck / racecheck / initcheck / synccheck):

    compute-sanitizer --tool memcheck  python tools/sanitize_run.py
    compute-sanitizer --tool racecheck python tools/sanitize_run.py

Covers the fast paths and, through the test knobs, the rare exact paths (Bloom slow path, CMS redo / spill /
direct paths).  Results are still compared with the oracle so that a sanitizer-only behaviour change shows.
No torch import: the sanitizer then only sees this library's kernels.
"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    import oracle_ffi
    from kamino_b200 import ffi, synth

    orc = oracle_ffi.load()
    species = synth.bacterial(3, seed=11, n_prot=60, mean_len=150.0)
    b = synth.stage(species)
    k, scheme = 13, ffi.SR6
    h = orc.cmsa_new()
    want_new = orc.count_batch(h, b.residues, b.rec_off, b.sp_rec_off, k, scheme, threads=4)
    want = np.array(orc.cmsa_table(h))
    orc.cmsa_free(h)
    min_needed = orc.min_needed(0.85, len(species))
    knob_sets = [{}, {"KMN_BLOOM_RUN_CAP": "8"}, {"KMN_CMS_FORCE": "1"}, {"KMN_CMS_FORCE": "2"}, {"KMN_CMS_SLOT": "4"},
                 {"KMN_CMS_CAP": "8", "KMN_CMS_SLOT": "2"}, {"KMN_H2D_CHUNKS": "1"}]
    only = os.environ.get("SANITIZE_ONLY")
    if only is not None:
        knob_sets = [knob_sets[int(only)]]
    for knobs in knob_sets:
        for kk in ("KMN_BLOOM_RUN_CAP", "KMN_CMS_FORCE", "KMN_CMS_SLOT", "KMN_CMS_CAP", "KMN_H2D_CHUNKS"):
            os.environ.pop(kk, None)
        os.environ.update(knobs)
        with ffi.Context(k, scheme, device=0) as ctx:
            bid, got_new = ctx.count_batch(b.residues, b.rec_off, b.sp_rec_off)
            ctx.finalize_table()
            assert got_new.tolist() == want_new.tolist(), knobs
            assert np.array_equal(ctx.table_snapshot(), want), knobs
            if knobs:
                print("ok", knobs, flush=True)
                continue
            ctx.scan_batch(bid, min_needed)
            cms = orc.cms_from_table(want)
            ws, we, _ = orc.scan_species(cms, b.residues, b.rec_off, 0, int(b.sp_rec_off[1]), k, scheme, min_needed)
            wp = orc.pair_species(cms, b.residues, b.rec_off, 0, int(b.sp_rec_off[1]), k, scheme, min_needed, 35)
            orc.cms_free(cms)
            gs, ge = ctx.fetch_hits(bid, 0)
            assert np.array_equal(gs, ws) and np.array_equal(ge, we)
            ctx.pair_batch(bid, 35)
            assert np.array_equal(ctx.fetch_pairs(bid, 0), wp)
            ctx.fetch_occupancy(bid, 1)
            # grouping on a few thousand synthetic observations
            rng = np.random.default_rng(3)
            n_obs = 5000
            left = rng.integers(0, 300, n_obs).astype(np.uint64)
            right = rng.integers(0, 3, n_obs).astype(np.uint64)
            sid = np.sort(rng.integers(0, 40, n_obs)).astype(np.uint32)
            length = rng.integers(27, 31, n_obs).astype(np.uint32)
            perm, groups = ctx.group_observations(left, right, sid, length, 3)
            assert int(groups["count"].sum()) == n_obs and sorted(perm.tolist()) == list(range(n_obs))
            # --nj counts, whole matrix and a row block
            m = rng.integers(0, 21, size=(130, 777), dtype=np.uint8)
            wv, wm = orc.pair_counts(m, threads=4)
            gv, gm = ctx.pair_counts(m)
            iu = np.triu_indices(130, 1)
            assert np.array_equal(gv[iu], wv[iu]) and np.array_equal(gm[iu], wm[iu])
            rv, rm = ctx.pair_counts_rows(m, 64, 130)
            assert np.array_equal(rv, gv[64:]) and np.array_equal(rm, gm[64:])
        print("ok", knobs, flush=True)
    print("sanitize_run: all paths equal the oracle")


if __name__ == "__main__":
    main()
