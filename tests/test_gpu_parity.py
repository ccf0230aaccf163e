"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.
Bit-exact: table, per-species new-k-mer counts, occupancies, start/end hit lists, pair counts."""
import numpy as np
import pytest

from kamino_b200 import synth

pytestmark = pytest.mark.gpu
DAYHOFF6, SR6, KGB6 = 0, 1, 2


@pytest.fixture(scope="module")
def ffi():
    from kamino_b200 import build, ffi as _ffi
    build.build()
    return _ffi


def _batch_from_records(species_records):
    """species_records: list (species) of list (records) of bytes."""
    res, rec_off, sp = [], [0], [0]
    for recs in species_records:
        for r in recs:
            res.append(np.frombuffer(r, dtype=np.uint8))
            rec_off.append(rec_off[-1] + len(r))
        sp.append(len(rec_off) - 1)
    residues = np.concatenate(res) if res else np.zeros(0, np.uint8)
    return residues, np.array(rec_off, np.uint64), np.array(sp, np.uint64)


def _check_pass1_and_pass2(ffi, oracle, residues, rec_off, sp, k, scheme, min_needed, scan_species=None,
                           threads=8, count=None):
    n_sp = len(sp) - 1
    h = oracle.cmsa_new()
    try:
        want_new = oracle.count_batch(h, residues, rec_off, sp, k, scheme, threads=threads)
        want_table = oracle.cmsa_table(h)
        with ffi.Context(k, scheme) as ctx:
            bid, got_new = count(ctx) if count else ctx.count_batch(residues, rec_off, sp)
            assert got_new.tolist() == want_new.tolist()
            ctx.finalize_table()
            got_table = ctx.table_snapshot()
            assert np.array_equal(got_table, want_table), "count-min table differs from the oracle"
            ctx.scan_batch(bid, min_needed)
            cms = oracle.cms_from_table(want_table)
            try:
                for s in (range(n_sp) if scan_species is None else scan_species):
                    ws, we, wocc = oracle.scan_species(cms, residues, rec_off, int(sp[s]), int(sp[s + 1]), k,
                                                       scheme, min_needed)
                    gs, ge = ctx.fetch_hits(bid, s)
                    gocc = ctx.fetch_occupancy(bid, s)
                    assert np.array_equal(gocc, wocc[:len(gocc)]), f"occupancy differs (species {s})"
                    assert np.array_equal(gs, ws), f"start candidates differ (species {s})"
                    assert np.array_equal(ge, we), f"end anchors differ (species {s})"
                # anchor selection + bubble pairing on the device (group_extraction.rs:151-242)
                for lm in (35, 3):
                    ctx.pair_batch(bid, lm)
                    for s in (range(n_sp) if scan_species is None else scan_species):
                        wp = oracle.pair_species(cms, residues, rec_off, int(sp[s]), int(sp[s + 1]), k, scheme,
                                                 min_needed, lm)
                        assert np.array_equal(ctx.fetch_pairs(bid, s), wp), f"anchor pairs differ (species {s}, l={lm})"
            finally:
                oracle.cms_free(cms)
        return want_new, got_table   # owned copy, verified equal to the oracle's (freed below)
    finally:
        oracle.cmsa_free(h)


def _load_3diff(oracle, golden_dir):
    names = ["A.fas.gz", "B.fas", "C.fas", "D1.fas", "D2.fas", "E1.fas", "E2.fas", "E3.fas"]
    recs = []
    for n in names:
        r, o = oracle.read_fasta(golden_dir / "test_3diff" / "input" / n)
        recs.append([r[int(o[i]):int(o[i + 1])].tobytes() for i in range(len(o) - 1)])
    return _batch_from_records(recs)


@pytest.mark.parametrize("scheme", [DAYHOFF6, SR6, KGB6])
def test_3diff_golden_inputs(ffi, oracle, golden_dir, scheme):
    residues, rec_off, sp = _load_3diff(oracle, golden_dir)
    new, table = _check_pass1_and_pass2(ffi, oracle, residues, rec_off, sp, 13, scheme, 7)
    assert new.tolist() == [84] * 8
    if scheme == DAYHOFF6:  # SURVEY.md §8c known answers
        nz = table[table != 0]
        assert nz.size == 440 and int(nz.sum()) == 2688


def test_3diff_known_anchors(ffi, oracle, golden_dir):
    residues, rec_off, sp = _load_3diff(oracle, golden_dir)
    with ffi.Context(13, DAYHOFF6) as ctx:
        bid, _ = ctx.count_batch(residues, rec_off, sp)
        ctx.finalize_table()
        ns, ne = ctx.scan_batch(bid, 7)
        starts, ends = ctx.fetch_hits(bid, 0)
        assert [(int(h["start"]), int(h["occ"]), int(h["kmer"])) for h in starts] == [(17, 8, 313726741273)]
        assert [(int(h["start"]), int(h["occ"]), int(h["kmer"])) for h in ends] == [(36, 8, 78946387236)]


@pytest.mark.parametrize("k,scheme", [(13, SR6), (1, SR6), (2, DAYHOFF6), (7, KGB6), (16, SR6), (21, DAYHOFF6)])
def test_synthetic_small(ffi, oracle, k, scheme):
    species = synth.bacterial(5, seed=11 + k, n_prot=300, mean_len=200.0)
    b = synth.stage(species)
    _check_pass1_and_pass2(ffi, oracle, b.residues, b.rec_off, b.sp_rec_off, k, scheme,
                           oracle.min_needed(0.85, 5))


def test_edge_cases_ragged(ffi, oracle):
    recs = [
        [b"", b"ACDEFGHIKLMNPQRSTVWY" * 3, b"", b""],                      # empty records around one
        [b"acdefghiklmnpqrstvwy\nACDEFGHIKLMNPQRSTVWYACDEFG\r\nHIKL \t MNPQ\x0cRSTVWY"],  # case + whitespace
        [],                                                                 # species without records
        [b"ACDEFGHIKLMN", b"ACDEFGHIKLMNP", b"ACDEFGHIKLMNPQ"],            # k-1, k, k+1 residues
        [b"ACDEFGHIKLMNPQRSTVWYACDEFGHXIKLMNPQRSTVWYACDEFGHIKLMNPQRSTVWY" * 2],  # X resets
        [b"ACD*EFG-HIK.LMNBPQRZSTVJWYUACDOEFG\x0bHIKLMNPQRSTVWY1ACDEFGHIKLMNPQRSTVWY\xffAC\x80DE"],
        [b"\n\n\n", b"   ", b"X", b"XXXXXXXXXXXXXXXXXXXXXXXXXXXX"],        # nothing countable
        [b"MVLFFQILGFALFIFWLLLIARVVVEFIRSFSRDWHPTGITVVVLEIIMSITDPPVKLLRRLIPQLTIGAVRFDLSIMVLLLVAFIGMQLAFGAAA"] * 3,
        [b"A" * 5000],                                                     # one k-mer repeated
    ]
    residues, rec_off, sp = _batch_from_records(recs)
    for k, scheme in [(13, SR6), (3, KGB6), (21, DAYHOFF6)]:
        _check_pass1_and_pass2(ffi, oracle, residues, rec_off, sp, k, scheme, 1)
        _check_pass1_and_pass2(ffi, oracle, residues, rec_off, sp, k, scheme, 2)


def test_empty_batches(ffi, oracle):
    with ffi.Context(13, SR6) as ctx:
        bid, new = ctx.count_batch(np.zeros(0, np.uint8), np.zeros(1, np.uint64), np.zeros(1, np.uint64))
        assert new.size == 0
        bid, new = ctx.count_batch(np.zeros(0, np.uint8), np.zeros(3, np.uint64), np.array([0, 1, 1, 2], np.uint64))
        assert new.tolist() == [0, 0, 0]
        ctx.finalize_table()
        assert not ctx.table_snapshot().any()
        assert ctx.scan_batch(bid, 1) == (0, 0)
        s, e = ctx.fetch_hits(bid, 1)
        assert len(s) == 0 and len(e) == 0


def test_argument_errors(ffi):
    with pytest.raises(ffi.KaminoError):
        ffi.Context(22, SR6)
    with pytest.raises(ffi.KaminoError):
        ffi.Context(0, SR6)
    with ffi.Context(13, SR6) as ctx:
        with pytest.raises(ffi.KaminoError):
            ctx.count_batch(np.zeros(4, np.uint8), np.array([0, 5, 3], np.uint64), np.array([0, 2], np.uint64))
        with pytest.raises(ffi.KaminoError):
            ctx.scan_batch(0, 1)   # no such batch / table not finalized
        with pytest.raises(ffi.KaminoError):
            ctx.table_snapshot()


def test_many_small_species_cross_groups(ffi, oracle):
    """More species than first-touch slots, species boundaries inside 16-residue chunks."""
    rng = np.random.default_rng(3)
    aa = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", np.uint8)
    recs = []
    for s in range(53):
        n_rec = int(rng.integers(0, 4))
        recs.append([aa[rng.integers(0, 20, int(rng.integers(0, 70)))].tobytes() for _ in range(n_rec)])
    residues, rec_off, sp = _batch_from_records(recs)
    _check_pass1_and_pass2(ffi, oracle, residues, rec_off, sp, 5, SR6, 2)


def test_bloom_order_dependence_large_species(ffi, oracle):
    """One 12 M-residue species: ~24 M probes into 2^25 bits, so a large share of the occurrences
    are Bloom false positives whose outcome depends on file order (proba_filter.rs:59-74)."""
    rng = np.random.default_rng(7)
    aa = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", np.uint8)
    seq = aa[rng.integers(0, 20, 12_000_000)]
    lens = np.full(12_000, 1000)
    b = synth.stage([synth.Species(seq, lens), synth.Species(seq[:2_000_000].copy(), lens[:2000])])
    new, _ = _check_pass1_and_pass2(ffi, oracle, b.residues, b.rec_off, b.sp_rec_off, 13, SR6, 2,
                                    scan_species=[1])
    n_kmers = 12_000 * (1000 - 12)
    assert new[0] < n_kmers * 0.97, "expected a visible Bloom false-positive rate"


def test_cms_saturation_and_repeat_calls(ffi, oracle):
    """70 000 species holding the same single k-mer: cells saturate at 65 535 (proba_filter.rs:152-156)."""
    one = b"ACDEFGHIKLMNP"
    recs = [[one] for _ in range(35_000)]
    residues, rec_off, sp = _batch_from_records(recs)
    key = int(oracle.roll_kmers(one, 13, SR6)[0])
    idx = oracle.cms_indices(key)
    with ffi.Context(13, SR6) as ctx:
        ctx.count_batch(residues, rec_off, sp, want_counts=False)
        ctx.finalize_table()
        t = ctx.table_snapshot()
        assert [int(t[r * (1 << 26) + idx[r]]) for r in range(4)] == [35_000] * 4
        ctx.reset_table()
        for _ in range(2):
            ctx.count_batch(residues, rec_off, sp, want_counts=False)
        ctx.finalize_table()
        t = ctx.table_snapshot()
        assert [int(t[r * (1 << 26) + idx[r]]) for r in range(4)] == [65_535] * 4
        assert int(t.astype(np.uint64).sum()) == 4 * 65_535


def test_staged_recount_is_idempotent_after_reset(ffi, oracle):
    species = synth.bacterial(3, seed=5, n_prot=200, mean_len=150.0)
    b = synth.stage(species)
    with ffi.Context(13, SR6) as ctx:
        bid = ctx.stage_batch(b.residues, b.rec_off, b.sp_rec_off)
        a = ctx.count_staged(bid, b.n_sp)
        ctx.finalize_table()
        t1 = ctx.table_snapshot()
        ctx.reset_table()
        a2 = ctx.count_staged(bid, b.n_sp)
        ctx.finalize_table()
        t2 = ctx.table_snapshot()
        assert a.tolist() == a2.tolist() and np.array_equal(t1, t2)
        # linearity of the sketch: counting twice doubles every (unsaturated) cell
        ctx.reset_table()
        ctx.count_staged(bid)
        ctx.count_staged(bid)
        ctx.finalize_table()
        assert np.array_equal(ctx.table_snapshot(), (t1.astype(np.uint32) * 2).astype(np.uint16))


def test_table_load_roundtrip(ffi, oracle):
    rng = np.random.default_rng(9)
    table = np.zeros(4 << 26, np.uint16)
    table[rng.integers(0, table.size, 100_000)] = rng.integers(1, 65535, 100_000)
    with ffi.Context(13, SR6) as ctx:
        ctx.table_load(table)
        assert np.array_equal(ctx.table_snapshot(), table)


@pytest.mark.parametrize("n,L", [(2, 1), (7, 131), (70, 1000), (130, 77), (65, 4097)])
def test_pair_counts(ffi, oracle, n, L):
    rng = np.random.default_rng(n * 1000 + L)
    m = rng.integers(0, 20, size=(n, L), dtype=np.uint8)
    m[rng.random((n, L)) < 0.07] = 20
    m[1:, :] = np.where(rng.random((n - 1, L)) < 0.8, m[0], m[1:, :])   # related rows
    wv, wm = oracle.pair_counts(m, threads=4)
    with ffi.Context(13, SR6) as ctx:
        gv, gm = ctx.pair_counts(m)
    iu = np.triu_indices(n, 1)
    assert np.array_equal(gv[iu], wv[iu]) and np.array_equal(gm[iu], wm[iu])


def test_pair_counts_row_blocks(ffi, oracle):
    """kmn_pair_counts_rows: every row block equals the same rows of the whole matrix (SURVEY 8e, K6 sharding)."""
    from kamino_b200 import sharding
    rng = np.random.default_rng(99)
    n, L = 333, 517
    m = rng.integers(0, 21, size=(n, L), dtype=np.uint8)
    wv, wm = oracle.pair_counts(m, threads=4)
    up = np.triu(np.ones((n, n), bool), 1)
    with ffi.Context(13, SR6) as ctx:
        for world in (1, 2, 5):
            for b0, b1 in sharding.pair_row_blocks(n, world):
                gv, gm = ctx.pair_counts_rows(m, b0, b1)
                assert gv.shape == (b1 - b0, n)
                assert np.array_equal(gv, np.where(up, wv, 0)[b0:b1]) and np.array_equal(gm, np.where(up, wm, 0)[b0:b1])
        # counts left on the device for a collective: same numbers behind kmn_pair_counts_device
        import torch
        from kamino_b200.sharding import _DevArray
        assert ctx.pair_counts_rows(m, 64, 192, fetch=False) is None
        pv, pm, rows, cols = ctx.pair_counts_device()
        assert (rows, cols) == (128, n)
        dv = torch.as_tensor(_DevArray(pv, (rows, cols), "<i4"), device="cuda").cpu().numpy().view(np.uint32)
        dm = torch.as_tensor(_DevArray(pm, (rows, cols), "<i4"), device="cuda").cpu().numpy().view(np.uint32)
        assert np.array_equal(dv, np.where(up, wv, 0)[64:192]) and np.array_equal(dm, np.where(up, wm, 0)[64:192])
        with pytest.raises(ffi.KaminoError):
            ctx.pair_counts_rows(m, 10, 64)        # not tile-aligned
        with pytest.raises(ffi.KaminoError):
            ctx.pair_counts_rows(m, 64, 400)       # past the matrix


@pytest.mark.parametrize("knobs", [
    {"KMN_CMS_FORCE": "1"},                      # every bucket redone with u32 counters
    {"KMN_CMS_FORCE": "2"},                      # direct CAS path only
    {"KMN_CMS_CAP": "8", "KMN_CMS_SLOT": "2"},   # spill list overflows -> fallback flag -> direct path
    {"KMN_CMS_SLOT": "1"},                       # almost every staged entry spills
    {"KMN_CMS_SLOT": "1", "KMN_CMS_FORCE": "1"},
])
def test_cms_rare_paths(ffi, oracle, knobs, monkeypatch):
    """The exact rare paths of the count-min build (cms_kernels.cuh) against the oracle."""
    for key, v in knobs.items():
        monkeypatch.setenv(key, v)
    species = synth.bacterial(6, seed=21, n_prot=400, mean_len=220.0)
    b = synth.stage(species)
    _, once = _check_pass1_and_pass2(ffi, oracle, b.residues, b.rec_off, b.sp_rec_off, 13, SR6, oracle.min_needed(0.85, 6),
                                     scan_species=[0])
    # kmn_reset_table zeroes lazily: stale cells must neither survive a reset + count nor a bare reset
    with ffi.Context(13, SR6) as ctx:
        ctx.table_load(np.full(4 << 26, 12345, np.uint16))
        ctx.reset_table()
        bid = ctx.stage_batch(b.residues, b.rec_off, b.sp_rec_off)
        ctx.count_staged(bid)
        ctx.finalize_table()
        assert np.array_equal(ctx.table_snapshot(), once)
        ctx.reset_table()
        ctx.finalize_table()
        assert not ctx.table_snapshot().any()
    # twice into the same table: merge on top of non-zero cells
    h = oracle.cmsa_new()
    try:
        for _ in range(2):
            oracle.count_batch(h, b.residues, b.rec_off, b.sp_rec_off, 13, SR6, threads=8)
        want = oracle.cmsa_table(h)
        with ffi.Context(13, SR6) as ctx:
            bid = ctx.stage_batch(b.residues, b.rec_off, b.sp_rec_off)
            ctx.count_staged(bid)
            ctx.count_staged(bid)
            ctx.finalize_table()
            assert np.array_equal(ctx.table_snapshot(), want)
    finally:
        oracle.cmsa_free(h)


def _repeat_heavy_species(rng):
    """Tandem repeats (period 1..7), a duplicated gene family and random proteins: many positions repeat
    a k-mer seen a few residues earlier, which stresses the same-step conflict list of the Bloom stage."""
    aa = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", np.uint8)
    recs = []
    for period in (1, 2, 3, 5, 7):
        unit = aa[rng.integers(0, 20, period)].tobytes()
        recs.append(unit * (4000 // period))
    gene = aa[rng.integers(0, 20, 400)].tobytes()
    recs += [gene] * 30
    recs += [aa[rng.integers(0, 20, int(rng.integers(30, 900)))].tobytes() for _ in range(200)]
    rng.shuffle(recs)
    return recs


def test_bloom_repeats_and_conflicts(ffi, oracle):
    rng = np.random.default_rng(17)
    residues, rec_off, sp = _batch_from_records([_repeat_heavy_species(rng) for _ in range(3)])
    for k in (13, 4):
        _check_pass1_and_pass2(ffi, oracle, residues, rec_off, sp, k, SR6, 2, scan_species=[0])


@pytest.mark.parametrize("cap", ["1", "16", "80"])
def test_bloom_slow_path(ffi, oracle, cap, monkeypatch):
    """Run slots too small for the probes of a tile: the species take bloom_slow_kernel (global bitset)."""
    monkeypatch.setenv("KMN_BLOOM_RUN_CAP", cap)
    rng = np.random.default_rng(23)
    species = synth.bacterial(4, seed=31, n_prot=300, mean_len=250.0)
    b = synth.stage(species)
    _check_pass1_and_pass2(ffi, oracle, b.residues, b.rec_off, b.sp_rec_off, 13, SR6, oracle.min_needed(0.85, 4),
                           scan_species=[1])
    residues, rec_off, sp = _batch_from_records([_repeat_heavy_species(rng) for _ in range(2)])
    _check_pass1_and_pass2(ffi, oracle, residues, rec_off, sp, 13, KGB6, 2, scan_species=[0])


def test_bloom_natural_overflow_takes_slow_path(ffi, oracle):
    """A period-5 tandem repeat of 40 000 residues: 5 k-mers, 8 000 occurrences each, far more than a run
    slot holds, next to ordinary species in the same batch."""
    rng = np.random.default_rng(29)
    aa = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", np.uint8)
    rep = [(b"ACDEW" * 8000)] + [aa[rng.integers(0, 20, 300)].tobytes() for _ in range(50)]
    normal = [aa[rng.integers(0, 20, int(rng.integers(50, 600)))].tobytes() for _ in range(300)]
    residues, rec_off, sp = _batch_from_records([normal, rep, normal[:100]])
    new, _ = _check_pass1_and_pass2(ffi, oracle, residues, rec_off, sp, 13, SR6, 2)
    assert new[1] > 0


def test_frame_tile_boundaries_and_empty_records(ffi, oracle):
    """Record boundaries exactly on 4096- and 16384-byte offsets (the frame tiles), empty records on both sides of
    them, and thousands of empty records inside one tile (frame_kernel then reads its boundaries from global memory
    and stores bytes directly)."""
    rng = np.random.default_rng(41)
    aa = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", np.uint8)

    def prot(n):
        return aa[rng.integers(0, 20, n)].tobytes()

    sp0 = [prot(4096), b"", b"", prot(4096 - 7), b"", prot(7), b"", prot(8192), b"\n" * 4096, prot(100)]
    sp1 = [b""] * 3000 + [prot(500)] + [b""] * 2500 + [prot(4096 * 2 - 500)] + [b""]
    sp2 = [prot(30) + b"\n" + prot(30) + b" \t" + prot(4036 - 63)] + [b"", prot(4096)]
    residues, rec_off, sp = _batch_from_records([sp0, sp1, [], sp2])
    for k in (13, 5):
        _check_pass1_and_pass2(ffi, oracle, residues, rec_off, sp, k, SR6, 1)


def test_frame_whitespace_runs_and_chunk_edges(ffi, oracle):
    """frame_kernel places the kept bytes of a 16-byte chunk run by run: chunks with several holes (single
    residues between line breaks), CRLF, whitespace runs longer than a chunk and longer than a thread's 64 bytes,
    invalid letters, records that start / end on 16-, 64- and 16384-byte edges, the whole thing across many tiles."""
    rng = np.random.default_rng(43)
    aa = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWYXBZ*acdefghiklmnpqrstvwy", np.uint8)
    ws = [b"\n", b"\r\n", b" ", b"\t", b"\n\n\n", b" " * 17, b"\n" * 70, b"\x0c"]

    def noisy(n_res, p_ws):
        out = bytearray()
        while n_res > 0:
            run = int(min(n_res, rng.integers(1, 90)))
            out += aa[rng.integers(0, len(aa), run)].tobytes()
            n_res -= run
            if rng.random() < p_ws:
                out += ws[int(rng.integers(0, len(ws)))]
        return bytes(out)

    def exact(n):   # exactly n raw bytes, line breaks every 60
        body = bytearray(aa[rng.integers(0, 20, n)].tobytes())
        body[60::61] = b"\n" * len(body[60::61])
        return bytes(body)

    sp0 = [noisy(int(rng.integers(0, 700)), 0.9) for _ in range(400)]
    sp1 = [exact(16), exact(48), exact(64), b"", exact(64), exact(128 - 1), exact(1), exact(16384 - 321), exact(16384), b"",
           exact(16384 + 16), b"A\nC\nD\nE\nF\nG\nH\nI\n" * 40, b"\r\n".join([b"ACDEFGHIKLMNPQ"] * 300)]
    sp2 = [noisy(50000, 0.5), b"\n" * 20000, noisy(3000, 1.0), b" "]
    residues, rec_off, sp = _batch_from_records([sp0, sp1, sp2])
    for k in (13, 4):
        _check_pass1_and_pass2(ffi, oracle, residues, rec_off, sp, k, SR6, 1)


def _host_threads():
    import os
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


def _check_pass2_species(ffi, oracle, ctx, bid, b, table, species_ids, k, scheme, min_needed, length_middle=35):
    """Hits, occupancy and anchor pairs of some species of a scanned + paired batch against the oracle."""
    cms = oracle.cms_from_table(table)
    try:
        for s in species_ids:
            r0, r1 = int(b.sp_rec_off[s]), int(b.sp_rec_off[s + 1])
            ws, we, wocc = oracle.scan_species(cms, b.residues, b.rec_off, r0, r1, k, scheme, min_needed)
            gs, ge = ctx.fetch_hits(bid, s)
            gocc = ctx.fetch_occupancy(bid, s)
            assert np.array_equal(gocc, wocc[:len(gocc)]), f"occupancy differs (species {s})"
            assert np.array_equal(gs, ws), f"start candidates differ (species {s})"
            assert np.array_equal(ge, we), f"end anchors differ (species {s})"
            wp = oracle.pair_species(cms, b.residues, b.rec_off, r0, r1, k, scheme, min_needed, length_middle)
            assert np.array_equal(ctx.fetch_pairs(bid, s), wp), f"anchor pairs differ (species {s})"
    finally:
        oracle.cms_free(cms)


def test_full_size_c1_equals_oracle(ffi, oracle):
    """BASELINE.json configs[1] at FULL size (100 synthetic bacterial proteomes, 1.2e8 residues): the 512 MiB
    count-min table, the per-species new-k-mer counts, and pass 2 (occupancy, boundary hits, anchor pairs) of
    five species are bit-identical to the oracle's (extract_groups stages 1-4, group_extraction.rs:527-608)."""
    species = synth.bacterial(100)
    b = synth.stage(species)
    k, scheme = 13, SR6
    min_needed = oracle.min_needed(0.85, 100)
    assert min_needed == 85
    h = oracle.cmsa_new()
    try:
        want_new = oracle.count_batch(h, b.residues, b.rec_off, b.sp_rec_off, k, scheme, threads=_host_threads())
        want = oracle.cmsa_table(h)
        with ffi.Context(k, scheme) as ctx:
            bid, got_new = ctx.count_batch(b.residues, b.rec_off, b.sp_rec_off)
            ctx.finalize_table()
            got = ctx.table_snapshot()
            assert got_new.tolist() == want_new.tolist(), "per-species new k-mer counts differ from the oracle"
            assert np.array_equal(got, want), "count-min table differs from the oracle at full size"
            assert int(got.astype(np.uint64).sum()) == 4 * int(want_new.sum())       # nothing saturates at C1
            ctx.scan_batch(bid, min_needed)
            ctx.pair_batch(bid, 35)
            _check_pass2_species(ffi, oracle, ctx, bid, b, want, (0, 1, 37, 63, 99), k, scheme, min_needed)
    finally:
        oracle.cmsa_free(h)


def test_c3_shape_eukaryotic_multi_batch_equals_oracle(ffi, oracle):
    """BASELINE.json configs[3] shape: eukaryotic proteomes (20 k proteins x ~500 aa, ~1e7 residues each: the
    Bloom filter ends a proteome 15-25 % full, so false positives matter, proba_filter.rs:59-74) counted in
    THREE batches that accumulate into one table (apply pass: overwrite, then read-modify-write twice).
    24 species, 2.4e8 residues: table, counts and pass 2 of four species equal the oracle's."""
    species = synth.eukaryotic(24)
    k, scheme = 13, SR6
    min_needed = oracle.min_needed(0.85, len(species))
    cuts = [0, 9, 17, 24]
    batches = [synth.stage(species[a:z]) for a, z in zip(cuts[:-1], cuts[1:])]
    whole = synth.stage(species)
    h = oracle.cmsa_new()
    try:
        want_new = oracle.count_batch(h, whole.residues, whole.rec_off, whole.sp_rec_off, k, scheme, threads=_host_threads())
        want = oracle.cmsa_table(h)
        # the regime this test is about: a visible share of the occurrences is filtered by the Bloom filter
        assert int(want_new.sum()) < 0.93 * whole.n_residues
        with ffi.Context(k, scheme) as ctx:
            ids, got_new = [], []
            for bt in batches:
                bid, n = ctx.count_batch(bt.residues, bt.rec_off, bt.sp_rec_off)
                ids.append(bid)
                got_new.append(n)
            ctx.finalize_table()
            got = ctx.table_snapshot()
            assert np.concatenate(got_new).tolist() == want_new.tolist()
            assert np.array_equal(got, want), "count-min table differs from the oracle (3 accumulating batches)"
            for bid, bt, sp_ids in ((ids[0], batches[0], (0, 8)), (ids[2], batches[2], (3, 6))):
                ctx.scan_batch(bid, min_needed)
                ctx.pair_batch(bid, 35)
                _check_pass2_species(ffi, oracle, ctx, bid, bt, want, sp_ids, k, scheme, min_needed)
    finally:
        oracle.cmsa_free(h)


def test_full_size_properties_100_proteomes(ffi, monkeypatch):
    """BASELINE.json configs[1] at full size (100 synthetic bacterial proteomes, 1.2e8 residues), checked
    through size-independent properties of the path instead of the (slow) oracle: the table is invariant
    under the order of the species and under the split into batches, per-species counts travel with their
    species, counting a batch twice doubles every unsaturated cell, and pass 2 on it is deterministic."""
    species = synth.bacterial(100)
    b = synth.stage(species)
    order = np.arange(99, -1, -1)
    rev = synth.stage([species[i] for i in order])
    half_a, half_b = synth.stage(species[:37]), synth.stage(species[37:])
    with ffi.Context(13, SR6) as ctx:
        bid, new = ctx.count_batch(b.residues, b.rec_off, b.sp_rec_off)
        ctx.finalize_table()
        t_all = ctx.table_snapshot()
        hits = ctx.scan_batch(bid, 85)
        pairs = ctx.pair_batch(bid, 35)
        p7 = ctx.fetch_pairs(bid, 7).copy()
        ctx.release_batches()
        assert int(new.sum()) > 0.9 * b.n_residues * 0.95 and hits[0] > 0 and pairs > 0

        ctx.reset_table()
        bid, new_rev = ctx.count_batch(rev.residues, rev.rec_off, rev.sp_rec_off)
        ctx.finalize_table()
        assert np.array_equal(ctx.table_snapshot(), t_all), "table depends on the species order"
        assert np.array_equal(new_rev, new[order]), "per-species counts do not follow their species"
        assert ctx.scan_batch(bid, 85) == hits
        assert ctx.pair_batch(bid, 35) == pairs
        assert np.array_equal(ctx.fetch_pairs(bid, 99 - 7), p7), "pass 2 of a species depends on its neighbours"
        ctx.release_batches()

        ctx.reset_table()
        _, na = ctx.count_batch(half_a.residues, half_a.rec_off, half_a.sp_rec_off)
        _, nb = ctx.count_batch(half_b.residues, half_b.rec_off, half_b.sp_rec_off)
        ctx.finalize_table()
        assert np.array_equal(ctx.table_snapshot(), t_all), "table depends on the batch split"
        assert np.array_equal(np.concatenate([na, nb]), new)
        ctx.release_batches()

        ctx.reset_table()
        bid = ctx.stage_batch(b.residues, b.rec_off, b.sp_rec_off)
        ctx.count_staged(bid)
        ctx.count_staged(bid)
        ctx.finalize_table()
        assert np.array_equal(ctx.table_snapshot(), np.minimum(t_all.astype(np.uint32) * 2, 65535).astype(np.uint16))
    # the copy / compute pipeline of kmn_count_batch: 1 and 16 species ranges give the same table as the default 5
    for chunks in ("1", "16"):
        monkeypatch.setenv("KMN_H2D_CHUNKS", chunks)
        with ffi.Context(13, SR6) as ctx:
            _, n2 = ctx.count_batch(b.residues, b.rec_off, b.sp_rec_off)
            ctx.finalize_table()
            assert np.array_equal(ctx.table_snapshot(), t_all) and np.array_equal(n2, new), f"KMN_H2D_CHUNKS={chunks}"


def _group_reference(left, right, sid, length, min_needed, k):
    """merge_species_extraction (group_extraction.rs:335-357) + select_unique_best_supported_length
    (group_sorting.rs:90-143) restated with numpy / python containers."""
    order = np.lexsort((np.arange(left.size), right, left))        # stable: ties keep emission order
    keys = np.stack([left[order], right[order]], axis=1)
    heads = np.ones(left.size, bool)
    heads[1:] = np.any(keys[1:] != keys[:-1], axis=1)
    firsts = np.flatnonzero(heads)
    out = []
    for g, i0 in enumerate(firsts):
        i1 = firsts[g + 1] if g + 1 < firsts.size else left.size
        by_len = {}
        for o in order[i0:i1]:
            by_len.setdefault(int(length[o]), set()).add(int(sid[o]))
        best_len, best_c, tie = 0, 0, False
        for l, sp in by_len.items():          # first-appearance order
            if len(sp) > best_c:
                best_len, best_c, tie = l, len(sp), False
            elif len(sp) == best_c:
                tie = True
        ok = (not tie) and best_c >= min_needed and best_len >= 2 * k + 1
        out.append((int(i0), int(i1 - i0), best_len, best_c, 1 if ok else 0))
    return order.astype(np.uint32), out


@pytest.mark.parametrize("n,n_keys,seed", [(1, 1, 0), (5000, 40, 1), (200_000, 3000, 2), (70_000, 70_000, 3)])
def test_group_observations(ffi, oracle, n, n_keys, seed):
    rng = np.random.default_rng(seed)
    k = 13
    pool_l = rng.integers(0, 1 << 39, n_keys, dtype=np.uint64)
    pool_r = rng.integers(0, 1 << 39, n_keys, dtype=np.uint64)
    pool_r[::7] = pool_r[0]                       # many pairs share one anchor
    pick = rng.integers(0, n_keys, n)
    sid = np.sort(rng.integers(0, 60, n)).astype(np.uint32)        # emission order is species-major
    length = (27 + rng.choice([0, 0, 0, 1, 5, 30], n)).astype(np.uint32)
    length[rng.random(n) < 0.02] = 20                               # shorter than 2k+1
    left, right = pool_l[pick], pool_r[pick]
    with ffi.Context(k, SR6) as ctx:
        perm, groups = ctx.group_observations(left, right, sid, length, 3)
    want_perm, want = _group_reference(left, right, sid, length, 3, k)
    assert np.array_equal(perm, want_perm)
    assert len(groups) == len(want)
    # the oracle's own merge + select_unique_best_supported_length (group_sorting.rs:90-143)
    og, osel = oracle.group_observations(left, right, sid, length, 60, 3, k)
    assert len(og) == len(groups)
    first = groups["first"].astype(np.int64)
    assert np.array_equal(left[perm[first]], og["left"]) and np.array_equal(right[perm[first]], og["right"])
    assert np.array_equal(groups["count"], og["count"])
    assert np.array_equal(groups["flags"] & ffi.GROUP_VALID, og["ok"])
    ok = og["ok"] == 1
    assert np.array_equal(groups["best_len"][ok], og["best_len"][ok]) and np.array_equal(groups["coverage"][ok], og["coverage"][ok])
    # selected observations (block length == best length, emission order inside the group)
    got_sel = []
    for g in groups[ok]:
        idx = perm[int(g["first"]):int(g["first"]) + int(g["count"])]
        got_sel.append(idx[length[idx] == g["best_len"]])
    got_sel = np.concatenate(got_sel) if got_sel else np.zeros(0, np.int64)
    assert np.array_equal(got_sel.astype(np.uint64), osel)
    for g, w in zip(groups, want):
        assert not (int(g["flags"]) & ffi.GROUP_HOST)
        assert (int(g["first"]), int(g["count"])) == w[:2]
        assert (int(g["flags"]) & ffi.GROUP_VALID) == w[4]
        if w[4]:
            assert (int(g["best_len"]), int(g["coverage"])) == w[2:4]


def test_group_observations_host_fallback_flags(ffi):
    """More distinct block lengths than the device tracks, and a run that is not species-major."""
    n = 40
    left = np.full(n, 5, np.uint64)
    right = np.full(n, 9, np.uint64)
    sid = np.arange(n, dtype=np.uint32)
    length = (30 + np.arange(n)).astype(np.uint32)
    left[30:] = 6
    length[30:] = 31
    sid[30:] = np.array([3, 4, 5, 2, 6, 7, 8, 9, 10, 11], np.uint32)     # 2 after 5: not species-major
    with ffi.Context(13, SR6) as ctx:
        perm, groups = ctx.group_observations(left, right, sid, length, 1)
    assert np.array_equal(perm, np.arange(n))
    assert [int(g["count"]) for g in groups] == [30, 10]
    assert all(int(g["flags"]) & ffi.GROUP_HOST for g in groups)


# ---- FASTA framing on the device (kmn_count_fasta): the oracle parses the same files on the CPU ----------

def _fasta_inputs(oracle, tmp_path, files):
    """files: raw bytes of each FASTA file.  Returns the oracle's (residues, rec_off, sp_rec_off), the bytes
    handed to kmn_count_fasta (a line feed appended to files that lack one) and their file offsets."""
    recs, parts, file_off = [], [], [0]
    for i, data in enumerate(files):
        p = tmp_path / f"f{i}.faa"
        p.write_bytes(data)
        r, o = oracle.read_fasta(p)
        recs.append([r[int(o[j]):int(o[j + 1])].tobytes() for j in range(len(o) - 1)])
        if data and not data.endswith(b"\n"):
            data += b"\n"
        parts.append(data)
        file_off.append(file_off[-1] + len(data))
    residues, rec_off, sp = _batch_from_records(recs)
    return residues, rec_off, sp, b"".join(parts), np.asarray(file_off, np.uint64), recs


def _fasta_text(sp, width=60, eol=b"\n", header=lambda pid: f">p{pid} some protein {pid} [Synth sp]".encode()):
    out, pos = [], 0
    for pid, ln in enumerate(sp.lens):
        out.append(header(pid) + eol)
        s = sp.seq[pos:pos + ln].tobytes()
        pos += int(ln)
        if width:
            out += [s[i:i + width] + eol for i in range(0, len(s), width)]
        else:
            out.append(s + eol)
    return b"".join(out)


def _adversarial_fasta_files():
    sps = synth.bacterial(5, seed=41, n_prot=90, mean_len=170.0)
    big = synth.bacterial(1, seed=42, n_prot=3, mean_len=9000.0)[0]
    files = [
        _fasta_text(sps[0]),                                                        # plain 60-column FASTA
        b"\n\r\n\n" + _fasta_text(sps[1], eol=b"\r\n"),                            # CRLF, leading empty lines
        _fasta_text(sps[2], width=0)[:-1],                                          # one line per sequence, no final line feed
        _fasta_text(sps[3], width=37, header=lambda pid: b">" + b"h>" * (pid % 7) + b"x" * (13 * pid % 5000)),   # '>' inside long headers
        b"\n\n\n",                                                                 # empty lines only: no record
        b"",                                                                        # empty file
        b">a\n>b\n\n>c desc\nACDEFGHIKLMNPQRSTVWY>ACDEFGHIKLMNPQRSTVWYACDEFGHIK\n  \t\n>d\nacdefghiklmnpqrstvwyACDEFGHIKLMNP\n>e",  # empty records, '>' inside a sequence, blanks, lower case, header at EOF
        _fasta_text(big, width=0),                                                  # lines far longer than a 4096-byte tile
        _fasta_text(sps[4]).lower(),
    ]
    return files


def test_count_fasta_matches_host_parsed_batch(ffi, oracle, tmp_path):
    files = _adversarial_fasta_files()
    residues, rec_off, sp, data, file_off, recs = _fasta_inputs(oracle, tmp_path, files)
    n_files = len(files)
    seen = {}

    def count(ctx):
        bid, new, n_rec = ctx.count_fasta(data, file_off)
        hdr, sp_rec = ctx.fetch_records(bid, n_files)
        seen.update(n_rec=n_rec, hdr=hdr, sp_rec=sp_rec)
        return bid, new

    _check_pass1_and_pass2(ffi, oracle, residues, rec_off, sp, 13, SR6, oracle.min_needed(0.6, n_files), count=count)
    assert seen["n_rec"] == len(rec_off) - 1 and seen["sp_rec"].tolist() == sp.tolist()
    # every record's '>' is where a host parser finds it
    want_hdr = []
    for f in range(n_files):
        chunk = data[int(file_off[f]):int(file_off[f + 1])]
        pos = 0
        for line in chunk.split(b"\n"):
            if line.startswith(b">"):
                want_hdr.append(int(file_off[f]) + pos)
            pos += len(line) + 1
    assert seen["hdr"].tolist() == want_hdr


@pytest.mark.parametrize("scheme", [DAYHOFF6, SR6])
def test_count_fasta_3diff_golden_files(ffi, oracle, golden_dir, tmp_path, scheme):
    import gzip
    names = ["A.fas.gz", "B.fas", "C.fas", "D1.fas", "D2.fas", "E1.fas", "E2.fas", "E3.fas"]
    files = []
    for n in names:
        raw = (golden_dir / "test_3diff" / "input" / n).read_bytes()
        files.append(gzip.decompress(raw) if n.endswith(".gz") else raw)
    residues, rec_off, sp, data, file_off, _ = _fasta_inputs(oracle, tmp_path, files)
    new, _ = _check_pass1_and_pass2(ffi, oracle, residues, rec_off, sp, 13, scheme, 7,
                                    count=lambda ctx: ctx.count_fasta(data, file_off)[:2])
    assert new.tolist() == [84] * 8


def test_count_fasta_rejects_bad_input(ffi):
    with ffi.Context(13, SR6) as ctx:
        good = b">a\nACDEFGHIKLMNPQRSTVWY\n"
        for bad, what in [(b"ACDE\n>a\nACDE\n", "expected '>'"), (b"\r>a\nACDE\n", "expected '>'"),
                          (b" \n>a\nACDE\n", "expected '>'"), (b">a\nACDE", "line feed")]:
            data = good + bad
            with pytest.raises(ffi.KaminoError, match=what):
                ctx.count_fasta(data, np.array([0, len(good), len(data)], np.uint64))
        # the context stays usable and nothing was counted
        bid, new, n_rec = ctx.count_fasta(good, np.array([0, len(good)], np.uint64))
        assert n_rec == 1 and new.tolist() == [8]
        ctx.finalize_table()
        assert int(ctx.table_snapshot().astype(np.int64).sum()) == 4 * 8


def test_count_fasta_full_size(ffi, oracle):
    """100 proteomes as FASTA text: same table and counts as the host-parsed batch (property at full size)."""
    species = synth.bacterial(100)
    b = synth.stage(species)
    parts, file_off = [], [0]
    for sp_ in species:
        t = _fasta_text(sp_)
        parts.append(t)
        file_off.append(file_off[-1] + len(t))
    data = np.frombuffer(b"".join(parts), np.uint8)
    with ffi.Context(13, SR6) as ctx:
        _, new_a = ctx.count_batch(b.residues, b.rec_off, b.sp_rec_off)
        ctx.finalize_table()
        table_a = ctx.table_snapshot()
        ctx.reset_table()
        ctx.release_batches()
        bid, new_b, n_rec = ctx.count_fasta(data, np.asarray(file_off, np.uint64))
        ctx.finalize_table()
        assert n_rec == len(b.rec_off) - 1
        assert new_a.tolist() == new_b.tolist()
        assert np.array_equal(ctx.table_snapshot(), table_a)


def test_group_batches_reads_the_device_pairs_in_place(ffi, oracle):
    """kmn_group_batches (observations = the anchor pairs kmn_pair_batch left on the device, two batches) against
    kmn_group_observations on the fetched pairs and against the oracle's merge + length selection
    (group_extraction.rs:335-357, group_sorting.rs:90-143)."""
    species = synth.species_set(12, 400, 250.0, seed=5)
    k, scheme = 13, SR6
    min_needed = oracle.min_needed(0.6, len(species))
    a, b = synth.stage(species[:7]), synth.stage(species[7:])
    with ffi.Context(k, scheme) as ctx:
        ids = []
        for bt in (a, b):
            bid, _ = ctx.count_batch(bt.residues, bt.rec_off, bt.sp_rec_off)
            ids.append(bid)
        ctx.finalize_table()
        pairs, sids = [], []
        for bid, bt, base in ((ids[0], a, 0), (ids[1], b, 7)):
            ctx.scan_batch(bid, min_needed)
            ctx.pair_batch(bid, 35)
            for s in range(bt.n_sp):
                p = ctx.fetch_pairs(bid, s)
                pairs.append(p)
                sids.append(np.full(len(p), base + s, np.uint32))
        pairs, sids = np.concatenate(pairs), np.concatenate(sids)
        assert len(pairs) > 1000
        perm_d, groups_d = ctx.group_batches(ids, min_needed, sid_base=[0, 7], full=True)
        perm_h, groups_h = ctx.group_observations(pairs["left"], pairs["right"], sids, pairs["len"], min_needed)
    assert np.array_equal(perm_d, perm_h)
    assert np.array_equal(groups_d, groups_h)
    og, osel = oracle.group_observations(pairs["left"], pairs["right"], sids, pairs["len"], len(species), min_needed, k)
    assert len(og) == len(groups_d)
    assert np.array_equal(groups_d["count"], og["count"])
    assert np.array_equal(groups_d["flags"] & ffi.GROUP_VALID, og["ok"])
    ok = og["ok"] == 1
    assert ok.any(), "expected some accepted groups"
    assert np.array_equal(groups_d["best_len"][ok], og["best_len"][ok])
    assert np.array_equal(groups_d["coverage"][ok], og["coverage"][ok])


def test_table_checksum_matches_the_snapshot(ffi, oracle):
    """kmn_table_checksum (row sums + position-sensitive hash, reduced on the device) against numpy on the snapshot."""
    b = synth.stage(synth.bacterial(3, seed=77, n_prot=200, mean_len=150.0))
    with ffi.Context(13, SR6) as ctx:
        ctx.reset_table()
        rows0, h0 = ctx.table_checksum()                     # a reset table that nothing was counted into: all zero
        assert rows0.tolist() == [0, 0, 0, 0] and h0 == 0
        _, new = ctx.count_batch(b.residues, b.rec_off, b.sp_rec_off)
        rows_live, h_live = ctx.table_checksum()             # before finalize: the live table
        ctx.finalize_table()
        rows, h = ctx.table_checksum()
        t = ctx.table_snapshot()
    assert (rows_live.tolist(), h_live) == (rows.tolist(), h)
    assert rows.tolist() == [int(new.sum())] * 4
    assert rows.tolist() == [int(t[r << 26:(r + 1) << 26].astype(np.uint64).sum()) for r in range(4)]
    nz = np.flatnonzero(t)
    want = 0
    for i in nz.tolist():
        want = (want + int(t[i]) * oracle.mix64(i + 0x51)) & 0xFFFFFFFFFFFFFFFF
    assert h == want


@pytest.mark.gpu
def test_pass2_narrow_table_matches_oracle_and_wide_scan(ffi, oracle, monkeypatch):
    """Pass 2 reads a one-byte-per-cell copy of the table (4 sweeps instead of 8) when every cell fits the byte window
    above min_needed - 1, and the u16 table otherwise; both give the oracle's hits, occupancies and pairs
    (group_extraction.rs:271-299: only comparisons with min_needed and between shared k-mers matter)."""
    rng = np.random.default_rng(5)
    aa = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)
    core = [aa[rng.integers(0, 20, 120)].tobytes() for _ in range(3)]
    recs = []
    for s in range(300):   # 300 small species: the core proteins' k-mers reach cells of ~300
        own = [aa[rng.integers(0, 20, 150)].tobytes() for _ in range(2)]
        mutated = bytearray(core[s % 3])
        if s % 7 == 0:
            mutated[40 + s % 30] = ord("W")
        recs.append([core[0], own[0], bytes(mutated), core[2] if s % 2 else own[1]])
    residues, rec_off, sp = _batch_from_records(recs)
    k, scheme = 13, SR6
    some = [0, 1, 7, 150, 299]
    # min_needed = 1: base 0, largest cell ~300 > 255 -> the u16 table; 60 / 255: the narrow table
    for min_needed in (1, 60, 255):
        _check_pass1_and_pass2(ffi, oracle, residues, rec_off, sp, k, scheme, min_needed, scan_species=some)
    # the narrow path is really taken, and equals the wide one on every species
    def scan(env_wide):
        if env_wide:
            monkeypatch.setenv("KMN_SCAN_WIDE", "1")
        else:
            monkeypatch.delenv("KMN_SCAN_WIDE", raising=False)
        with ffi.Context(k, scheme) as ctx:
            bid, _ = ctx.count_batch(residues, rec_off, sp)
            ctx.finalize_table()
            before = ctx.launch_count()
            ctx.scan_batch(bid, 60)
            launches = ctx.launch_count() - before
            hits = [ctx.fetch_hits(bid, s) for s in range(len(recs))]
        return launches, hits
    ln, hn = scan(False)
    lw, hw = scan(True)
    monkeypatch.delenv("KMN_SCAN_WIDE", raising=False)
    assert lw - ln == 2, (ln, lw)   # 8 sweeps vs table max + narrow copy + 4 sweeps
    for (sn, en), (sw, ew) in zip(hn, hw):
        assert np.array_equal(sn, sw) and np.array_equal(en, ew)


@pytest.mark.gpu
def test_host_batch_range_schedules_match_staged_count(ffi, monkeypatch):
    """kmn_count_batch cuts a host batch into species ranges whose copies overlap the kernels (record offsets travel
    in slices with them); batches over 160 MB take the long schedule (a ramp up to 48 MB ranges).  Whatever the
    schedule, table and per-species counts equal those of the same batch staged first and counted in one range
    (count_species_kmers is order-free across species: group_extraction.rs:558-567)."""
    species = synth.bacterial(150, seed=77)          # ~180 MB: the long schedule
    b = synth.stage(species)
    assert b.residues.size > (160 << 20)
    pin, h_res = ffi.PinnedBuffer.copy_of(b.residues)
    pin_rec, h_rec = ffi.PinnedBuffer.copy_of(b.rec_off)
    try:
        with ffi.Context(13, SR6) as ctx:
            bid = ctx.stage_batch(b.residues, b.rec_off, b.sp_rec_off)
            want_new = ctx.count_staged(bid, b.n_sp)
            ctx.finalize_table()
            want_rows, want_hash = ctx.table_checksum()
            assert [int(x) for x in want_rows] == [int(want_new.sum())] * 4
        for chunks in (None, "1", "3", "32"):          # long schedule; one range; 3 and 32 geometric ranges
            if chunks is None:
                monkeypatch.delenv("KMN_H2D_CHUNKS", raising=False)
            else:
                monkeypatch.setenv("KMN_H2D_CHUNKS", chunks)
            with ffi.Context(13, SR6) as ctx:
                # pinned residues + pinned offsets, then pageable ones
                for res, rec in ((h_res, h_rec), (b.residues, b.rec_off)):
                    ctx.reset_table()
                    _, new = ctx.count_batch(res, rec, b.sp_rec_off)
                    ctx.finalize_table()
                    rows, h = ctx.table_checksum()
                    assert new.tolist() == want_new.tolist(), f"per-species counts differ (KMN_H2D_CHUNKS={chunks})"
                    assert h == want_hash and [int(x) for x in rows] == [int(x) for x in want_rows], f"table differs (KMN_H2D_CHUNKS={chunks})"
                    ctx.release_batches()
    finally:
        monkeypatch.delenv("KMN_H2D_CHUNKS", raising=False)
        pin.free()
        pin_rec.free()


@pytest.mark.gpu
def test_stage_batch_async_pipeline_equals_oracle(ffi, oracle):
    """kmn_stage_batch_async + kmn_count_staged + kmn_release_batch, driven the way a multi-batch job drives them (the
    next batch is staged before the current one is counted, two device batches alternate): table and per-species
    counts equal the oracle's for the concatenation of the batches (AtomicCountMinSketch::increment is commutative,
    proba_filter.rs:146-166), and the staged batches can still be scanned."""
    batches = [synth.stage(synth.bacterial(3, seed=300 + i, n_prot=120 + 40 * i, mean_len=150.0)) for i in range(5)]
    pins = []
    host = []
    for b in batches:
        triple = []
        for a in (b.residues, b.rec_off, b.sp_rec_off):
            p, v = ffi.PinnedBuffer.copy_of(a)
            pins.append(p)
            triple.append(v)
        host.append(triple)
    k, scheme = 13, SR6
    h = oracle.cmsa_new()
    try:
        want_new = [oracle.count_batch(h, b.residues, b.rec_off, b.sp_rec_off, k, scheme, threads=4) for b in batches]
        want = oracle.cmsa_table(h)
        with ffi.Context(k, scheme) as ctx:
            got_new = []
            nxt = ctx.stage_batch_async(*host[0])
            for i in range(len(batches)):
                cur = nxt
                if i + 1 < len(batches):
                    nxt = ctx.stage_batch_async(*host[i + 1])
                got_new.append(ctx.count_staged(cur, batches[i].n_sp))
                if i != 2:
                    ctx.release_batch(cur)
                else:
                    kept = cur
            for g, w in zip(got_new, want_new):
                assert g.tolist() == w.tolist()
            ctx.finalize_table()
            assert np.array_equal(ctx.table_snapshot(), want), "table of the pipelined job differs from the oracle"
            # a batch that was kept can be scanned; a released id is an error; a staged batch nobody used can be released
            ns, ne = ctx.scan_batch(kept, 2)
            cms = oracle.cms_from_table(want)
            try:
                b = batches[2]
                ws, we, _ = oracle.scan_species(cms, b.residues, b.rec_off, 0, int(b.sp_rec_off[1]), k, scheme, 2)
                gs, ge = ctx.fetch_hits(kept, 0)
                assert np.array_equal(gs, ws) and np.array_equal(ge, we)
            finally:
                oracle.cms_free(cms)
            with pytest.raises(ffi.KaminoError):
                ctx.count_staged(0)
            unused = ctx.stage_batch_async(*host[1])
            ctx.release_batch(unused)
            with pytest.raises(ffi.KaminoError):
                ctx.stage_batch_async(batches[0].residues.astype(np.int32), host[0][1], host[0][2])
            # the order check of rec_off is deferred: an unsorted rec_off fails the first call that uses the batch
            bad_rec = host[0][1].copy()
            bad_rec[3], bad_rec[4] = bad_rec[4], bad_rec[3] + 0
            if bad_rec[3] >= bad_rec[4]:
                bad = ctx.stage_batch_async(host[0][0], bad_rec, host[0][2])
                ctx.reset_table()
                with pytest.raises(ffi.KaminoError, match="non-decreasing"):
                    ctx.count_staged(bad)
                ctx.release_batch(bad)
    finally:
        oracle.cmsa_free(h)
        for p in pins:
            p.free()
