"""ctypes wrapper of the CPU oracle (oracle/_build/libkamino_oracle.so) — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
LIB = ORACLE_DIR / "_build" / "libkamino_oracle.so"
CLI = ORACLE_DIR / "_build" / "kamino_oracle"
CMS_CELLS = 4 << 26

HIT_DTYPE = np.dtype(
    [("kmer", "<u8"), ("rec", "<u4"), ("seg_start", "<u4"), ("start", "<u4"), ("occ", "<u4")]
)
PAIR_DTYPE = np.dtype(
    [("left", "<u8"), ("right", "<u8"), ("rec", "<u4"), ("seg_start", "<u4"), ("left_start", "<u4"), ("len", "<u4")]
)


def build(force: bool = False) -> Path:
    srcs = list(ORACLE_DIR.glob("*.cpp")) + list(ORACLE_DIR.glob("*.hpp")) + [ORACLE_DIR / "Makefile"]
    stale = force or not LIB.exists() or not CLI.exists() or any(
        s.stat().st_mtime > LIB.stat().st_mtime for s in srcs)
    if stale:
        subprocess.run(["make", "-C", str(ORACLE_DIR), "-j8"], check=True, capture_output=True)
    return LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    def __init__(self, lib: C.CDLL):
        self.lib = lib
        vp = C.c_void_p
        lib.orc_last_error.restype = C.c_char_p
        lib.orc_recode_byte.argtypes = [C.c_uint8, C.c_int]
        lib.orc_recode_byte.restype = C.c_uint8
        lib.orc_mix64.argtypes = [C.c_uint64]
        lib.orc_mix64.restype = C.c_uint64
        lib.orc_k_mask.argtypes = [C.c_int]
        lib.orc_k_mask.restype = C.c_uint64
        lib.orc_min_needed.argtypes = [C.c_float, C.c_uint64]
        lib.orc_min_needed.restype = C.c_uint32
        lib.orc_roll_kmers.argtypes = [vp, C.c_uint64, C.c_int, C.c_int, vp, C.c_uint64]
        lib.orc_roll_kmers.restype = C.c_uint64
        lib.orc_bloom_new.argtypes = [C.c_uint64]
        lib.orc_bloom_new.restype = vp
        lib.orc_bloom_test_and_set.argtypes = [vp, C.c_uint64]
        lib.orc_bloom_free.argtypes = [vp]
        lib.orc_bloom_indices.argtypes = [C.c_uint64, vp, vp]
        lib.orc_cms_indices.argtypes = [C.c_uint64, vp]
        lib.orc_cmsa_new.restype = vp
        lib.orc_cmsa_free.argtypes = [vp]
        lib.orc_cmsa_increment.argtypes = [vp, C.c_uint64]
        lib.orc_cmsa_reset.argtypes = [vp]
        lib.orc_cmsa_table.argtypes = [vp]
        lib.orc_cmsa_table.restype = vp
        lib.orc_cmsa_snapshot.argtypes = [vp]
        lib.orc_cmsa_snapshot.restype = vp
        lib.orc_cms_from_table.argtypes = [vp]
        lib.orc_cms_from_table.restype = vp
        lib.orc_cms_table.argtypes = [vp]
        lib.orc_cms_table.restype = vp
        lib.orc_cms_estimate.argtypes = [vp, C.c_uint64]
        lib.orc_cms_estimate.restype = C.c_uint32
        lib.orc_cms_free.argtypes = [vp]
        lib.orc_count_batch.argtypes = [vp, vp, vp, C.c_uint64, vp, C.c_uint32, C.c_int, C.c_int, C.c_int, vp]
        lib.orc_scan_species.argtypes = [vp, vp, vp, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_uint32,
                                         vp, C.c_uint64, vp, vp, C.c_uint64, vp, vp]
        lib.orc_pair_species.argtypes = [vp, vp, vp, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_uint32, C.c_uint64,
                                         vp, C.c_uint64, vp]
        lib.orc_pair_counts.argtypes = [vp, C.c_uint32, C.c_uint64, vp, vp, C.c_int]
        lib.orc_aa_index.argtypes = [C.c_uint8]
        lib.orc_aa_index.restype = C.c_uint8
        lib.orc_f81_constant.restype = C.c_double
        lib.orc_lg_f81_from_counts.argtypes = [C.c_uint64, C.c_uint64]
        lib.orc_lg_f81_from_counts.restype = C.c_double
        lib.orc_nj_newick.argtypes = [vp, C.c_uint32, vp, vp, C.c_uint64]
        lib.orc_read_fasta.argtypes = [C.c_char_p, vp, C.c_uint64, vp, vp, C.c_uint64, vp]

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.orc_last_error().decode("utf-8", "replace"))

    # scalars
    def recode_byte(self, b, scheme): return self.lib.orc_recode_byte(b, scheme)
    def mix64(self, x): return self.lib.orc_mix64(x)
    def k_mask(self, k): return self.lib.orc_k_mask(k)
    def min_needed(self, f, n): return self.lib.orc_min_needed(f, n)

    def roll_kmers(self, seq: bytes, k: int, scheme: int) -> np.ndarray:
        a = np.frombuffer(seq, dtype=np.uint8)
        out = np.zeros(max(len(a), 1), dtype=np.uint64)
        n = self.lib.orc_roll_kmers(_p(a), len(a), k, scheme, _p(out), len(out))
        return out[:n]

    def bloom_indices(self, key):
        a, b = C.c_uint64(), C.c_uint64()
        self.lib.orc_bloom_indices(key, C.byref(a), C.byref(b))
        return a.value, b.value

    def cms_indices(self, key):
        out = np.zeros(4, dtype=np.uint64)
        self.lib.orc_cms_indices(key, _p(out))
        return [int(x) for x in out]

    # pass 1
    def cmsa_new(self):
        h = self.lib.orc_cmsa_new()
        if not h:
            raise MemoryError("oracle CMS allocation failed")
        return h

    def cmsa_free(self, h): self.lib.orc_cmsa_free(h)
    def cmsa_reset(self, h): self.lib.orc_cmsa_reset(h)
    def cmsa_increment(self, h, key): self.lib.orc_cmsa_increment(h, key)

    def cmsa_table(self, h) -> np.ndarray:
        """Zero-copy view of the atomic table (valid while the handle lives)."""
        p = self.lib.orc_cmsa_table(h)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint16)), shape=(CMS_CELLS,))

    def count_batch(self, h, residues, rec_off, sp_rec_off, k, scheme, threads=8) -> np.ndarray:
        residues = np.ascontiguousarray(residues, np.uint8)
        rec_off = np.ascontiguousarray(rec_off, np.uint64)
        sp_rec_off = np.ascontiguousarray(sp_rec_off, np.uint64)
        n_sp = len(sp_rec_off) - 1
        out = np.zeros(n_sp, dtype=np.uint64)
        self._check(self.lib.orc_count_batch(h, _p(residues), _p(rec_off), len(rec_off) - 1, _p(sp_rec_off),
                                             n_sp, k, scheme, threads, _p(out)))
        return out

    # pass 2
    def cms_from_table(self, table: np.ndarray):
        table = np.ascontiguousarray(table, np.uint16)
        assert table.size == CMS_CELLS
        h = self.lib.orc_cms_from_table(_p(table))
        if not h:
            raise MemoryError("oracle CMS allocation failed")
        return h

    def cms_free(self, h): self.lib.orc_cms_free(h)
    def cms_estimate(self, h, key): return self.lib.orc_cms_estimate(h, key)

    def scan_species(self, cms, residues, rec_off, rec_begin, rec_end, k, scheme, min_needed, want_occ=True):
        residues = np.ascontiguousarray(residues, np.uint8)
        rec_off = np.ascontiguousarray(rec_off, np.uint64)
        n_bytes = int(rec_off[rec_end] - rec_off[rec_begin])
        cap = max(n_bytes, 1)
        starts = np.zeros(cap, dtype=HIT_DTYPE)
        ends = np.zeros(cap, dtype=HIT_DTYPE)
        occ = np.zeros(cap, dtype=np.uint16) if want_occ else None
        ns, ne = C.c_uint64(), C.c_uint64()
        self._check(self.lib.orc_scan_species(cms, _p(residues), _p(rec_off), rec_begin, rec_end, k, scheme,
                                              min_needed, _p(starts), cap, C.byref(ns), _p(ends), cap,
                                              C.byref(ne), _p(occ)))
        return starts[:ns.value], ends[:ne.value], occ

    def pair_species(self, cms, residues, rec_off, rec_begin, rec_end, k, scheme, min_needed, length_middle):
        residues = np.ascontiguousarray(residues, np.uint8)
        rec_off = np.ascontiguousarray(rec_off, np.uint64)
        cap = max(int(rec_off[rec_end] - rec_off[rec_begin]), 1)
        out = np.zeros(cap, dtype=PAIR_DTYPE)
        n = C.c_uint64()
        self._check(self.lib.orc_pair_species(cms, _p(residues), _p(rec_off), rec_begin, rec_end, k, scheme,
                                              min_needed, length_middle, _p(out), cap, C.byref(n)))
        return out[:n.value]

    def group_observations(self, left, right, sid, length, n_species, min_needed, k):
        """merge_species_extraction + select_unique_best_supported_length (group_sorting.rs:90-143) on
        observations in emission order.  Returns (groups, selected): groups = structured array
        (left, right, count, ok, best_len, coverage) in ascending (left, right); selected = emission
        indices of the selected observations of the accepted groups, group after group."""
        left = np.ascontiguousarray(left, np.uint64); right = np.ascontiguousarray(right, np.uint64)
        sid = np.ascontiguousarray(sid, np.uint32); length = np.ascontiguousarray(length, np.uint32)
        n = left.size
        cap = max(n, 1)
        gl, gr = np.zeros(cap, np.uint64), np.zeros(cap, np.uint64)
        gc, gok, glen, gcov = (np.zeros(cap, np.uint32) for _ in range(4))
        sel = np.zeros(cap, np.uint64)
        ng, ns = C.c_uint64(), C.c_uint64()
        self.lib.orc_group_observations.argtypes = [C.c_uint64] + [C.c_void_p] * 4 + [C.c_uint64, C.c_uint32, C.c_int] + \
            [C.c_void_p] * 6 + [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        self._check(self.lib.orc_group_observations(n, _p(left), _p(right), _p(sid), _p(length), n_species, min_needed, k,
                                                    _p(gl), _p(gr), _p(gc), _p(gok), _p(glen), _p(gcov), cap,
                                                    C.byref(ng), _p(sel), C.byref(ns)))
        g = ng.value
        groups = np.zeros(g, dtype=[("left", "<u8"), ("right", "<u8"), ("count", "<u4"), ("ok", "<u4"),
                                    ("best_len", "<u4"), ("coverage", "<u4")])
        groups["left"], groups["right"], groups["count"] = gl[:g], gr[:g], gc[:g]
        groups["ok"], groups["best_len"], groups["coverage"] = gok[:g], glen[:g], gcov[:g]
        return groups, sel[:ns.value].copy()

    # --nj
    def pair_counts(self, mapped: np.ndarray, threads=8):
        mapped = np.ascontiguousarray(mapped, np.uint8)
        n, L = mapped.shape
        valid = np.zeros((n, n), dtype=np.uint32)
        mism = np.zeros((n, n), dtype=np.uint32)
        self._check(self.lib.orc_pair_counts(_p(mapped), n, L, _p(valid), _p(mism), threads))
        return valid, mism

    def aa_index(self, b): return self.lib.orc_aa_index(b)
    def f81_constant(self): return self.lib.orc_f81_constant()
    def lg_f81_from_counts(self, v, m): return self.lib.orc_lg_f81_from_counts(v, m)

    def nj_newick(self, names, dist: np.ndarray) -> str:
        n = len(names)
        arr = (C.c_char_p * n)(*[s.encode() for s in names])
        dist = np.ascontiguousarray(dist, np.float64)
        buf = C.create_string_buffer(64 * n + 1024 + sum(len(s) * 3 for s in names))
        self._check(self.lib.orc_nj_newick(arr, n, _p(dist), buf, len(buf)))
        return buf.value.decode()

    def self_test(self):
        """The reference's unit tests of the host logic on the oracle's functions (oracle/selftest.cpp)."""
        self.lib.orc_self_test.restype = C.c_int
        self._check(self.lib.orc_self_test())

    def read_fasta(self, path):
        nres, nrec = C.c_uint64(), C.c_uint64()
        self._check(self.lib.orc_read_fasta(str(path).encode(), None, 0, C.byref(nres), None, 0, C.byref(nrec)))
        res = np.zeros(nres.value, dtype=np.uint8)
        off = np.zeros(nrec.value + 1, dtype=np.uint64)
        self._check(self.lib.orc_read_fasta(str(path).encode(), _p(res), res.size, C.byref(nres), _p(off),
                                            off.size, C.byref(nrec)))
        return res, off

    def run_cli(self, args, cwd=None):
        return subprocess.run([str(CLI), *map(str, args)], cwd=cwd, capture_output=True, text=True)


_oracle = None


def load() -> Oracle:
    global _oracle
    if _oracle is None:
        build()
        _oracle = Oracle(C.CDLL(str(LIB)))
    return _oracle
