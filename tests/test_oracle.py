"""CPU tests: pin the oracle against the reference's golden fixtures and the derived known-answer
vectors of SURVEY.md §8c.  (The reference has no unit tests for recode.rs / proba_filter.rs.)"""
import shutil

import numpy as np
import pytest

DAYHOFF6, SR6, KGB6 = 0, 1, 2
GOLDEN_FILES = ["kamino_alignment.fas", "kamino_missing.tsv", "kamino_partitions.tsv", "kamino_NJ.tree"]


def _load_3diff(oracle, golden_dir):
    names = ["A.fas.gz", "B.fas", "C.fas", "D1.fas", "D2.fas", "E1.fas", "E2.fas", "E3.fas"]
    res, rec_off, sp = [], [0], [0]
    for n in names:
        r, o = oracle.read_fasta(golden_dir / "test_3diff" / "input" / n)
        res.append(r)
        rec_off.extend(int(x) + rec_off[-1] for x in o[1:])
        sp.append(len(rec_off) - 1)
    return np.concatenate(res), np.array(rec_off, np.uint64), np.array(sp, np.uint64)


# ---- reference goldens (tests/golden.rs:96-124 + the checked-in, never-compared NJ tree)
@pytest.mark.parametrize("mode,scheme", [("dir", "dayhoff6"), ("tsv", "kgb6")])
def test_golden_3diff(oracle, golden_dir, tmp_path, mode, scheme):
    src = golden_dir / "test_3diff"
    args = ["-i", src / "input"] if mode == "dir" else ["-I", src / "input.tsv"]
    r = oracle.run_cli([*args, "--length-middle", 35, "--recode", scheme, "--nj", "-t", 1,
                        "-o", tmp_path / "kamino"])
    assert r.returncode == 0, r.stderr
    for f in GOLDEN_FILES:
        assert (tmp_path / f).read_text() == (src / "expected" / f).read_text(), f


def test_golden_threads_independent(oracle, golden_dir, tmp_path):
    src = golden_dir / "test_3diff"
    r = oracle.run_cli(["-i", src / "input", "-l", 35, "-r", "dayhoff6", "--NJ", "-t", 4, "-o", tmp_path / "k"])
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "k_alignment.fas").read_text() == (src / "expected" / "kamino_alignment.fas").read_text()
    assert (tmp_path / "k_NJ.tree").read_text() == (src / "expected" / "kamino_NJ.tree").read_text()


def test_cli_validation(oracle, golden_dir, tmp_path):
    src = golden_dir / "test_3diff" / "input"
    for bad in (["-k", 22], ["-k", 0], ["-f", 0.5], ["-t", 0], ["-c", 14], ["--genomes"]):
        r = oracle.run_cli(["-i", src, "-o", tmp_path / "x", *bad])
        assert r.returncode != 0, bad


# ---- derived KATs
def test_mix64_kat(oracle):
    assert oracle.mix64(0) == 0
    assert oracle.mix64(1) == 0x5692161D100B05E5
    assert oracle.mix64(0x9E3779B97F4A7C15) == 0xE220A8397B1DCDAF  # SplitMix64 first output, seed 0


def test_first_kmer_kat(oracle):
    seq = b"MVLFFQILGFALF"
    assert oracle.roll_kmers(seq, 13, DAYHOFF6)[0] == 314292968037
    assert oracle.roll_kmers(seq, 13, SR6)[0] == 235909519901
    assert oracle.roll_kmers(seq, 13, KGB6)[0] == 183142778900


def test_sketch_indices_kat(oracle):
    key = 313726741273
    assert oracle.bloom_indices(key) == (23392436, 5644049)
    assert oracle.cms_indices(key) == [51456468, 20944087, 36763837, 61588028]


def test_recode_tables(oracle):
    classes = {
        DAYHOFF6: ["C", "AGPST", "DENQ", "HKR", "ILMV", "FWY"],
        SR6: ["APST", "DENG", "QKR", "MIVL", "WC", "FYH"],
        KGB6: ["AGPS", "DENQHKRT", "MIL", "W", "FY", "CV"],
    }
    for scheme, cls in classes.items():
        want = {}
        for code, letters in enumerate(cls):
            for ch in letters:
                want[ord(ch)] = code
                want[ord(ch.lower())] = code
        assert len(want) == 40
        for b in range(256):
            assert oracle.recode_byte(b, scheme) == want.get(b, 255), (scheme, b)


def test_rolling_resets_and_whitespace(oracle):
    k = 3
    a = oracle.roll_kmers(b"ACDEF", k, SR6)
    assert len(a) == 3
    # whitespace {SP,TAB,LF,FF,CR} is skipped without a reset; VT (0x0B) and 'X' reset
    assert list(oracle.roll_kmers(b"AC\nD \tE\x0c\rF", k, SR6)) == list(a)
    assert len(oracle.roll_kmers(b"AC\x0bDEF", k, SR6)) == 1
    assert len(oracle.roll_kmers(b"ACXDEF", k, SR6)) == 1
    assert list(oracle.roll_kmers(b"acdef", k, SR6)) == list(a)
    assert oracle.k_mask(21) == (1 << 63) - 1 and oracle.k_mask(13) == (1 << 39) - 1


def test_min_needed_f32(oracle):
    assert oracle.min_needed(0.85, 8) == 7
    assert oracle.min_needed(0.85, 100) == 85
    assert oracle.min_needed(0.85, 1000) == 850
    assert oracle.min_needed(0.85, 65535) == 55705
    assert oracle.min_needed(0.85, 0) == 1


def test_bloom_sequential_semantics(oracle):
    """test_and_set against a plain-Python restatement of proba_filter.rs:59-74."""
    rng = np.random.default_rng(1)
    keys = [int(x) for x in rng.integers(0, 1 << 39, 3000)]
    keys += keys[:500]
    mask = (1 << 10) - 1
    bits = set()
    h = oracle.lib.orc_bloom_new(1 << 10)
    try:
        for key in keys:
            h1 = oracle.mix64(key ^ 0x9E3779B97F4A7C15)
            h2 = oracle.mix64(key ^ 0xBF58476D1CE4E5B9) | 1
            idx = [h1 & mask, (h1 + h2) & mask]
            seen = all(i in bits for i in idx)
            bits.update(idx)
            assert bool(oracle.lib.orc_bloom_test_and_set(h, key)) == seen
    finally:
        oracle.lib.orc_bloom_free(h)


def test_cms_saturation(oracle):
    h = oracle.cmsa_new()
    try:
        key = 12345
        for _ in range(65540):
            oracle.cmsa_increment(h, key)
        t = oracle.cmsa_table(h)
        idx = oracle.cms_indices(key)
        for r in range(4):
            assert t[r * (1 << 26) + idx[r]] == 65535
    finally:
        oracle.cmsa_free(h)


def test_3diff_table_and_anchors(oracle, golden_dir):
    res, rec_off, sp = _load_3diff(oracle, golden_dir)
    h = oracle.cmsa_new()
    try:
        new = oracle.count_batch(h, res, rec_off, sp, 13, DAYHOFF6, threads=2)
        assert list(new) == [84] * 8
        t = oracle.cmsa_table(h)
        nz = t[t != 0]
        assert nz.size == 440 and int(nz.sum()) == 2688
        vals, cnts = np.unique(nz, return_counts=True)
        assert dict(zip(vals.tolist(), cnts.tolist())) == {1: 52, 3: 52, 4: 32, 5: 20, 7: 20, 8: 264}
        cms = oracle.cms_from_table(t)
        try:
            assert oracle.min_needed(0.85, 8) == 7
            starts, ends, occ = oracle.scan_species(cms, res, rec_off, 0, 1, 13, DAYHOFF6, 7)
            assert [(int(s["start"]), int(s["occ"])) for s in starts] == [(17, 8)]
            assert [(int(e["start"]), int(e["occ"])) for e in ends] == [(36, 8)]
            assert int(starts[0]["kmer"]) == 313726741273 and int(ends[0]["kmer"]) == 78946387236
            assert occ[:96].tolist()[:12] == [0] * 12 and occ[12] > 0
        finally:
            oracle.cms_free(cms)
    finally:
        oracle.cmsa_free(h)


def test_f81_kat(oracle):
    assert abs(oracle.f81_constant() - 0.940740913711) < 1e-12
    assert oracle.lg_f81_from_counts(9, 1) == 0.12568819053858765
    d = oracle.lg_f81_from_counts(9, 0)
    assert d == 0.0 and np.signbit(d)          # identical pair -> -0.0 (phylo.rs:105)
    d0 = oracle.lg_f81_from_counts(0, 0)
    assert d0 == 0.0 and not np.signbit(d0)    # no shared valid site -> +0.0 (phylo.rs:96-98)
    assert oracle.lg_f81_from_counts(10, 10) == -np.log(1.0 - 0.999999999)
    for i, ch in enumerate(b"ARNDCQEGHILKMFPSTWYV"):
        assert oracle.aa_index(ch) == i
    for ch in b"-XBZJUO*acd":
        assert oracle.aa_index(ch) == 20


def test_pair_counts_small(oracle):
    rng = np.random.default_rng(5)
    m = rng.integers(0, 21, size=(7, 131), dtype=np.uint8)
    v, mm = oracle.pair_counts(m, threads=2)
    for i in range(7):
        for j in range(i + 1, 7):
            ok = (m[i] < 20) & (m[j] < 20)
            assert v[i, j] == ok.sum() and mm[i, j] == (ok & (m[i] != m[j])).sum()


def test_golden_genomes_with_the_python_predictor(oracle, golden_dir, tmp_path):
    """tests/golden.rs:126-139 inside the oracle alone: oracle/protein_prediction.py writes the proteomes, the
    oracle CLI runs the pipeline, outputs equal the reference's expected files (pins the Python restatement)."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    from oracle import protein_prediction as pp
    src = golden_dir / "test_genomes"
    pred = tmp_path / "pred"
    pred.mkdir()
    for f in sorted((src / "input").iterdir()):
        name = f.name[:-3] if f.name.endswith(".gz") else f.name
        name = name.rsplit(".", 1)[0]
        (pred / f"{name}.faa").write_text(pp.predict_one_genome(f))
    r = oracle.run_cli(["-i", pred, "--length-middle", 50, "--recode", "sr6", "-t", 2, "-o", tmp_path / "kamino"])
    assert r.returncode == 0, r.stderr
    for f in ("kamino_alignment.fas", "kamino_missing.tsv", "kamino_partitions.tsv"):
        assert (tmp_path / f).read_bytes() == (src / "expected" / f).read_bytes(), f
    # known answers of the scoring arithmetic (hand-computed from the reference's formulas)
    assert pp.gc_model(pp.f32(0.5)) == (576, 400, 549)
    assert pp.length_score(270) == 31 and pp.calibrated_start_term(13, 600) == 6
    assert pp.translate_table11("GTGAAATAGNNN", True) == "MK*X"


def test_fasta_reader_agrees_with_the_python_restatement(oracle, tmp_path):
    """seq_io's record rules restated twice, independently (oracle/fasta_io.cpp and oracle/protein_prediction.py):
    same records on awkward files.  (The crate itself is not available here: SURVEY 8c, 'parity unpinned'.)"""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    from oracle import protein_prediction as pp
    texts = [
        b">a desc\nACDE\nFGH\n>b\nKLMN\n",
        b"\n\r\n>a\r\nAC\r\nDE\r\n>b x y\r\n\r\nFG\r\n",
        b">a\n>b\n\n>c d\nAC>DE\n \t\n>e",
        b">only header",
        b"",
        b"\n\n",
        b">x>y>z\nMKV\n>w\nLLL",
        b">a\nAC\n\n\n>b\n\nDE",
    ]
    for i, t in enumerate(texts):
        p = tmp_path / f"t{i}.fa"
        p.write_bytes(t)
        res, off = oracle.read_fasta(p)
        got = [res[int(off[j]):int(off[j + 1])].tobytes() for j in range(len(off) - 1)]
        want = [r.encode("latin-1") for r in pp.read_records(p)]
        assert got == want, (i, got, want)
    bad = tmp_path / "bad.fa"
    bad.write_bytes(b"ACDE\n>a\nAC\n")
    with pytest.raises(Exception):
        oracle.read_fasta(bad)
    with pytest.raises(ValueError):
        pp.read_records(bad)


def test_reference_unit_tests_on_the_oracle(oracle):
    """group_extraction.rs:666-748, group_sorting.rs:438-659 and group_filtering.rs:413-693 — the reference's own
    unit tests of the host logic — restated on the oracle's functions (oracle/selftest.cpp): 37 tests."""
    oracle.self_test()


def test_reference_unit_tests_on_the_python_predictor():
    """protein_prediction.rs:1152-1286 on oracle/protein_prediction.py."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    from oracle import protein_prediction as pp
    # lookup_tables_encode_bases_complements_starts_and_stops
    assert [pp.NT2.get(c) for c in "ACGTNa"] == [0, 1, 2, 3, None, None]
    assert [pp.COMPLEMENT.get(c, "N") for c in "ATCGan"] == ["T", "A", "G", "C", "t", "N"]
    assert {pp.codon_index(*c) for c in ("ATG", "GTG", "TTG")} == set(pp.START_BONUS) and pp.codon_index(*"ACG") not in pp.START_BONUS
    assert {pp.codon_index(*c) for c in ("TAA", "TAG", "TGA")} == set(pp.STOPS) and pp.codon_index(*"TGG") not in pp.STOPS
    gct, aaa = pp.codon_index(*"GCT"), pp.codon_index(*"AAA")

    def orf(a, b, strand, score, p5=False):
        return {"start": a, "end": b, "strand": strand, "scan_start": a, "scan_end": b, "p5": p5, "p3": False,
                "has_start": True, "score": score}
    # adaptive_model_training_uses_high_scoring_complete_orfs_and_background
    fwd, orfs = "", []
    for i in range(pp.ADAPTIVE_MIN_TRAIN_ORFS):
        orfs.append(orf(len(fwd), len(fwd) + 390, 1, 10_000 - i))
        fwd += "GCT" * 130
    for i in range(pp.ADAPTIVE_MIN_TRAIN_ORFS):
        orfs.append(orf(len(fwd), len(fwd) + 390, 1, i))
        fwd += "AAA" * 130
    model = pp.train_adaptive(orfs, fwd, "CCC" * (pp.ADAPTIVE_MIN_TRAIN_ORFS * 130))
    assert model is not None and model[gct] > 0 and model[aaa] < 0
    # adaptive_model_training_rejects_insufficient_or_partial_seed_orfs
    n = pp.ADAPTIVE_MIN_TRAIN_ORFS * pp.ADAPTIVE_TRAIN_MIN_AA
    assert pp.train_adaptive([orf(0, 360, 1, 0)] * (pp.ADAPTIVE_MIN_TRAIN_ORFS - 1), "ATG" * n, "CAT" * n) is None
    assert pp.train_adaptive([orf(0, 360, 1, 0, p5=True)] * pp.ADAPTIVE_MIN_TRAIN_ORFS, "ATG" * n, "CAT" * n) is None
    # adaptive_scores_are_clamped_and_added_per_strand
    m = [0] * 64
    m[gct], m[aaa] = 2200, -2200
    assert pp.adaptive_coding_score("GCT" * 34, m) == 0 and pp.adaptive_coding_score("GCT" * 35, m) == 18
    assert pp.adaptive_coding_score("AAA" * 35, m) == -18
    # accumulate_codon_counts_counts_only_complete_valid_codons
    counts = [0] * 64
    pp.codon_counts("ATGGCTNNNT", counts)
    assert counts[pp.codon_index(*"ATG")] == 1 and counts[gct] == 1 and sum(counts) == 2


def test_oracle_grouping_matches_a_numpy_restatement(oracle):
    """orc_group_observations (the K7 checker: merge_species_extraction + select_unique_best_supported_length,
    group_extraction.rs:335-357, group_sorting.rs:90-143) against an independent numpy/python restatement."""
    rng = np.random.default_rng(5)
    n, n_keys, k = 30_000, 900, 13
    pool_l = rng.integers(0, 1 << 39, n_keys, dtype=np.uint64)
    pool_r = rng.integers(0, 1 << 39, n_keys, dtype=np.uint64)
    pick = rng.integers(0, n_keys, n)
    sid = np.sort(rng.integers(0, 40, n)).astype(np.uint32)
    length = (27 + rng.choice([0, 0, 0, 1, 5, 30], n)).astype(np.uint32)
    length[rng.random(n) < 0.02] = 20
    left, right = pool_l[pick], pool_r[pick]
    groups, sel = oracle.group_observations(left, right, sid, length, 40, 3, k)
    order = np.lexsort((np.arange(n), right, left))
    heads = np.ones(n, bool)
    heads[1:] = (left[order][1:] != left[order][:-1]) | (right[order][1:] != right[order][:-1])
    firsts = np.flatnonzero(heads)
    assert len(groups) == firsts.size
    want_sel = []
    for g, i0 in enumerate(firsts):
        i1 = firsts[g + 1] if g + 1 < firsts.size else n
        idx = order[i0:i1]
        by_len = {}
        for o in idx:
            by_len.setdefault(int(length[o]), set()).add(int(sid[o]))
        best_len, best_c, tie = 0, 0, False
        for l, sp in by_len.items():
            if len(sp) > best_c:
                best_len, best_c, tie = l, len(sp), False
            elif len(sp) == best_c:
                tie = True
        ok = (not tie) and best_c >= 3 and best_len >= 2 * k + 1
        assert int(groups["count"][g]) == i1 - i0 and int(groups["ok"][g]) == int(ok)
        assert (int(groups["left"][g]), int(groups["right"][g])) == (int(left[idx[0]]), int(right[idx[0]]))
        if ok:
            assert (int(groups["best_len"][g]), int(groups["coverage"][g])) == (best_len, best_c)
            want_sel.append(idx[length[idx] == best_len])
    assert np.array_equal(np.concatenate(want_sel).astype(np.uint64), sel)
