"""CPU tests of two ARGUMENTS the CUDA path rests on, restated in numpy / on the oracle (no GPU, no product code):

* the k-mers of `kmn_device.cuh::load_chunk_kmers` — windows of the packed symbol string of a 16-code chunk and the two
  chunks before it, validity from the reset bits OR-ed over windows of k — equal the reference's roll / have recurrence
  (group_extraction.rs:371-388) at every position;
* pass 2 needs the estimates only where they are >= min_needed (`scan_kernels.cuh`, narrow table): replacing every table
  cell below min_needed by 0 leaves the oracle's start candidates, end anchors and anchor pairs unchanged
  (group_extraction.rs:271-299, :151-242).

The `-m gpu` tests check the kernels themselves against the oracle; these pin the reasoning."""
import numpy as np
import pytest

from kamino_b200 import synth   # numpy-only generator of related proteomes (no CUDA library behind it)

SR6 = 1
CHUNK = 16
INVALID = 0xFF


def _reference_kmers(codes: np.ndarray, k: int):
    """roll / have of count_species_kmers on a recoded stream (0..5, 0xFF = reset): (valid, kmer) per position."""
    mask = (1 << (3 * k)) - 1
    roll, have = 0, 0
    valid = np.zeros(codes.size, bool)
    kmer = np.zeros(codes.size, object)
    for i, c in enumerate(codes.tolist()):
        if c == INVALID:
            roll, have = 0, 0
            continue
        roll = ((roll << 3) | c) & mask
        have = min(have + 1, k)
        if have == k:
            valid[i] = True
            kmer[i] = roll
    return valid, kmer


def _pack_chunk(chunk):
    """pack_chunk_mask: symbol j at bits [3 (15 - j), +3) of P (0xFF packs as 7), bit j of the mask = code j is a reset."""
    p, m = 0, 0
    for j, c in enumerate(chunk):
        p = (p << 3) | (c & 7)
        if c == INVALID:
            m |= 1 << j
    return p, m


def _window_kmers(codes: np.ndarray, k: int):
    """load_chunk_kmers, chunk by chunk (a lane's two predecessors are the two chunks before it, whatever the warp)."""
    assert codes.size % CHUNK == 0
    mask = (1 << (3 * k)) - 1
    n_chunks = codes.size // CHUNK
    packed = [_pack_chunk(codes[c * CHUNK:(c + 1) * CHUNK].tolist()) for c in range(n_chunks)]
    valid = np.zeros(codes.size, bool)
    kmer = np.zeros(codes.size, object)
    for c in range(n_chunks):
        P, I = packed[c]
        P1, I1 = packed[c - 1] if c >= 1 else (0, 0x8000)      # before the buffer: the last symbol counts as a reset
        P2, I2 = packed[c - 2] if c >= 2 else (0, 0x8000)
        string = P | (P1 << 48) | ((P2 & 0xFFFFFFFF) << 96)    # only the low 32 bits of the chunk before the previous one
        acc = I2 | (I1 << 16) | (I << 32)
        w = 1
        while 2 * w <= k:
            acc |= acc << w
            w *= 2
        if k > w:
            acc |= acc << (k - w)
        acc &= (1 << 64) - 1                                    # the device works in 64 bits
        v = ~(acc >> 32) & 0xFFFF
        for j in range(CHUNK):
            kmer[c * CHUNK + j] = (string >> (3 * (15 - j))) & mask
            valid[c * CHUNK + j] = bool((v >> j) & 1)
    return valid, kmer


@pytest.mark.parametrize("k", [1, 2, 3, 7, 13, 16, 20, 21])
def test_window_kmers_equal_the_rolling_recurrence(k):
    rng = np.random.default_rng(100 + k)
    for p_reset in (0.0, 0.01, 0.08, 0.5):
        codes = rng.integers(0, 6, 64 * CHUNK).astype(np.uint8)
        codes[rng.random(codes.size) < p_reset] = INVALID
        if p_reset:
            codes[[0, 15, 16, 31, 32, 47]] = INVALID          # resets on chunk edges
            codes[200:200 + 3 * CHUNK] = rng.integers(0, 6, 3 * CHUNK)   # and a long clean stretch
        want_valid, want = _reference_kmers(codes, k)
        got_valid, got = _window_kmers(codes, k)
        assert np.array_equal(got_valid, want_valid), f"validity differs (k={k}, p={p_reset})"
        assert all(int(got[i]) == int(want[i]) for i in np.flatnonzero(want_valid)), f"k-mers differ (k={k}, p={p_reset})"


def test_pass2_needs_only_the_estimates_from_min_needed_up(oracle):
    """Zeroing every cell below min_needed changes no hit and no anchor pair: exactly what the narrow-table scan relies on
    (its occ[] is the estimate where >= min_needed and 0 elsewhere)."""
    b = synth.stage(synth.bacterial(9, seed=3, n_prot=60, mean_len=160.0))
    residues, rec_off, sp = b.residues, b.rec_off, b.sp_rec_off
    k = 13
    h = oracle.cmsa_new()
    try:
        oracle.count_batch(h, residues, rec_off, sp, k, SR6, threads=4)
        table = oracle.cmsa_table(h).copy()
    finally:
        oracle.cmsa_free(h)
    assert int(table.max()) >= 9
    for min_needed in (1, 3, 6, 8, 9):
        masked = np.where(table >= min_needed, table, 0).astype(np.uint16)
        full, cut = oracle.cms_from_table(table), oracle.cms_from_table(masked)
        try:
            some_hits = 0
            for s in range(len(sp) - 1):
                r0, r1 = int(sp[s]), int(sp[s + 1])
                ws, we, _ = oracle.scan_species(full, residues, rec_off, r0, r1, k, SR6, min_needed, want_occ=False)
                gs, ge, _ = oracle.scan_species(cut, residues, rec_off, r0, r1, k, SR6, min_needed, want_occ=False)
                assert np.array_equal(gs, ws) and np.array_equal(ge, we), f"hits differ (species {s}, min_needed {min_needed})"
                for lm in (35, 3):
                    wp = oracle.pair_species(full, residues, rec_off, r0, r1, k, SR6, min_needed, lm)
                    gp = oracle.pair_species(cut, residues, rec_off, r0, r1, k, SR6, min_needed, lm)
                    assert np.array_equal(gp, wp), f"pairs differ (species {s}, min_needed {min_needed}, l={lm})"
                some_hits += len(ws) + len(we)
            assert some_hits > 0 or min_needed == 9
        finally:
            oracle.cms_free(full)
            oracle.cms_free(cut)


def test_partition_slot_addresses_are_what_the_consumers_read():
    """The staging-store addresses of the two partition kernels, exhaustively over (bucket, rank):
    cms_partition_kernel writes entry index Tq ^ rank with Tq = ((idx >> 11) & 0x7FE0) | (bucket & ~0x7FE0) (one LOP3;
    idx < 2^26, bucket = idx >> 16) where cms_apply_kernel reads rank r of slot b at b * 32 + (r ^ (b & 31));
    bloom_partition_kernel<96> writes (stage + 384 b + 4 r) ^ 4 (b & 31) where bloom_bucket_kernel reads rank r of run b
    at word 96 b + (r ^ (b & 31)) (stage 128-byte aligned, slots a multiple of 128 bytes)."""
    rng = np.random.default_rng(1)
    b = np.arange(1024, dtype=np.uint64)[:, None]
    r = np.arange(32, dtype=np.uint64)[None, :]
    low = rng.integers(0, 1 << 16, (1024, 1), dtype=np.uint64)       # the entry's low 16 index bits must not matter
    idx = (b << np.uint64(16)) | low
    tq = ((idx >> np.uint64(11)) & np.uint64(0x7FE0)) | (b & np.uint64(~0x7FE0 & 0xFFFFFFFF))
    assert np.array_equal(tq ^ r, b * np.uint64(32) + (r ^ (b & np.uint64(31))))
    assert np.unique(tq ^ r).size == 1024 * 32                       # a bijection onto the staging buffer's 32768 entries
    for stage in (0x400, 0x12380, 0x30000):                          # 128-byte aligned shared-memory addresses
        bb = np.arange(256, dtype=np.uint64)[:, None]
        rr = np.arange(96, dtype=np.uint64)[None, :]
        got = (np.uint64(stage) + np.uint64(384) * bb + np.uint64(4) * rr) ^ (np.uint64(4) * (bb & np.uint64(31)))
        want = np.uint64(stage) + np.uint64(4) * (np.uint64(96) * bb + (rr ^ (bb & np.uint64(31))))
        assert np.array_equal(got, want)
