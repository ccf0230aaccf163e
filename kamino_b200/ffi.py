"""ctypes binding of libkamino_b200.so (C ABI in include/kamino_b200.h).

Thin plumbing only: numpy arrays in, numpy arrays out.  There is no CPU fallback — if the CUDA
library is missing or no GPU is present, every call raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import build as _build

DAYHOFF6, SR6, KGB6 = 0, 1, 2
SCHEMES = {"dayhoff6": DAYHOFF6, "sr6": SR6, "kgb6": KGB6}
CMS_CELLS = 4 << 26
STAGES = {"frame": 0, "bloom": 1, "cms": 2, "finalize": 3, "scan": 4, "pair_counts": 5,
          "cms_partition_kernel": 6, "bloom_bucket_kernel": 7, "pair_batch": 8, "group_observations": 9,
          "bloom_partition_kernel": 10, "cms_apply_kernel": 11, "h2d_gbs": 12}

HIT_DTYPE = np.dtype(
    [("kmer", "<u8"), ("rec", "<u4"), ("seg_start", "<u4"), ("start", "<u4"), ("occ", "<u4")]
)

PAIR_DTYPE = np.dtype(
    [("left", "<u8"), ("right", "<u8"), ("rec", "<u4"), ("seg_start", "<u4"), ("left_start", "<u4"), ("len", "<u4")]
)

GROUP_DTYPE = np.dtype(
    [("first", "<u4"), ("count", "<u4"), ("best_len", "<u4"), ("coverage", "<u4"), ("flags", "<u4")]
)
GROUP_VALID, GROUP_HOST = 1, 2

# every symbol declared in include/kamino_b200.h
ABI_SYMBOLS = [
    "kmn_last_error", "kmn_abi_version", "kmn_create", "kmn_destroy", "kmn_host_alloc", "kmn_host_free", "kmn_count_batch",
    "kmn_stage_batch", "kmn_stage_batch_async", "kmn_release_batch", "kmn_count_staged", "kmn_count_fasta", "kmn_fetch_records", "kmn_reset_table", "kmn_release_batches",
    "kmn_finalize_table", "kmn_table_snapshot", "kmn_table_load", "kmn_table_accum_device_ptr",
    "kmn_table_widen", "kmn_table_device_ptr", "kmn_exchange_export", "kmn_exchange_setup", "kmn_exchange_reduce", "kmn_exchange_setup_local", "kmn_exchange_reduce_all", "kmn_exchange_abort", "kmn_table_checksum", "kmn_synchronize", "kmn_stream", "kmn_scan_batch", "kmn_fetch_hits",
    "kmn_pair_batch", "kmn_fetch_pairs", "kmn_group_observations", "kmn_group_batches", "kmn_fetch_occupancy", "kmn_pair_counts", "kmn_pair_counts_rows", "kmn_pair_counts_device", "kmn_last_stage_ms", "kmn_launch_count",
]


class KaminoError(RuntimeError):
    pass


_lib = None


def library_path() -> Path:
    import os
    override = os.environ.get("KAMINO_B200_LIB")   # tuning builds of tools/ (same ABI, other -D constants)
    return Path(override) if override else _build.LIB_PATH


def load_library() -> C.CDLL:
    """Load libkamino_b200.so (must have been built: `python -m kamino_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not path.exists():
        raise KaminoError(
            f"{path} is missing: build it with `python -m kamino_b200.build` (no CPU fallback exists)"
        )
    lib = C.CDLL(str(path))
    u8p, u16p, u32p, u64p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint16, C.c_uint32, C.c_uint64))
    vp, vpp = C.c_void_p, C.POINTER(C.c_void_p)
    lib.kmn_last_error.restype = C.c_char_p
    lib.kmn_abi_version.restype = C.c_uint32
    lib.kmn_create.argtypes = [vpp, C.c_int, C.c_int, C.c_int]
    lib.kmn_destroy.argtypes = [vp]
    lib.kmn_destroy.restype = None
    lib.kmn_host_alloc.argtypes = [vpp, C.c_uint64]
    lib.kmn_host_free.argtypes = [vp]
    lib.kmn_host_free.restype = None
    lib.kmn_count_batch.argtypes = [vp, vp, vp, C.c_uint64, vp, C.c_uint32, vp, u32p]
    lib.kmn_stage_batch.argtypes = [vp, vp, vp, C.c_uint64, vp, C.c_uint32, u32p]
    lib.kmn_stage_batch_async.argtypes = [vp, vp, vp, C.c_uint64, vp, C.c_uint32, u32p]
    lib.kmn_release_batch.argtypes = [vp, C.c_uint32]
    lib.kmn_count_staged.argtypes = [vp, C.c_uint32, vp]
    lib.kmn_reset_table.argtypes = [vp]
    lib.kmn_release_batches.argtypes = [vp]
    lib.kmn_finalize_table.argtypes = [vp]
    lib.kmn_table_snapshot.argtypes = [vp, vp]
    lib.kmn_table_load.argtypes = [vp, vp]
    lib.kmn_table_accum_device_ptr.argtypes = [vp, vpp]
    lib.kmn_table_widen.argtypes = [vp]
    lib.kmn_table_device_ptr.argtypes = [vp, vpp]
    lib.kmn_exchange_export.argtypes = [vp, vp, vp]
    lib.kmn_exchange_setup.argtypes = [vp, C.c_int, C.c_int, vp, vp]
    lib.kmn_exchange_reduce.argtypes = [vp]
    lib.kmn_exchange_abort.argtypes = [vp]
    lib.kmn_exchange_setup_local.argtypes = [vpp, C.c_int]
    lib.kmn_exchange_reduce_all.argtypes = [vpp, C.c_int]
    lib.kmn_table_checksum.argtypes = [vp, vp, u64p]
    lib.kmn_synchronize.argtypes = [vp]
    lib.kmn_stream.argtypes = [vp, vpp]
    lib.kmn_scan_batch.argtypes = [vp, C.c_uint32, C.c_uint32, u64p, u64p]
    lib.kmn_fetch_hits.argtypes = [vp, C.c_uint32, C.c_uint32, vp, C.c_uint64, u64p, vp, C.c_uint64, u64p]
    lib.kmn_fetch_occupancy.argtypes = [vp, C.c_uint32, C.c_uint32, vp, C.c_uint64, u64p]
    lib.kmn_pair_batch.argtypes = [vp, C.c_uint32, C.c_uint32, u64p]
    lib.kmn_group_observations.argtypes = [vp, C.c_uint64, vp, vp, vp, vp, C.c_uint32, vp, vp, C.c_uint64, u64p]
    lib.kmn_fetch_pairs.argtypes = [vp, C.c_uint32, C.c_uint32, vp, C.c_uint64, u64p]
    lib.kmn_group_batches.argtypes = [vp, vp, C.c_uint32, vp, C.c_uint32, C.c_uint32, vp, C.c_uint64, vp, C.c_uint64, u64p, u64p]
    lib.kmn_pair_counts.argtypes = [vp, vp, C.c_uint32, C.c_uint64, vp, vp]
    lib.kmn_count_fasta.argtypes = [vp, vp, vp, C.c_uint32, vp, C.POINTER(C.c_uint32), u64p]
    lib.kmn_fetch_records.argtypes = [vp, C.c_uint32, vp, C.c_uint64, u64p, vp]
    lib.kmn_pair_counts_rows.argtypes = [vp, vp, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, vp, vp]
    lib.kmn_pair_counts_device.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.kmn_last_stage_ms.argtypes = [vp, C.c_int, C.POINTER(C.c_float)]
    lib.kmn_launch_count.argtypes = [vp]
    lib.kmn_launch_count.restype = C.c_uint64
    _lib = lib
    return lib


def _ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _as(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dtype)


def exchange_setup_local(contexts: list["Context"]):
    """Wire the exchange buffers of contexts that live in THIS process, one per device (kmn_exchange_setup_local)."""
    lib = load_library()
    arr = (C.c_void_p * len(contexts))(*[c._h for c in contexts])
    if lib.kmn_exchange_setup_local(arr, len(contexts)) != 0:
        raise KaminoError(lib.kmn_last_error().decode("utf-8", "replace"))


def exchange_reduce_all(contexts: list["Context"]):
    """One exchange over all the in-process contexts (kmn_exchange_reduce_all)."""
    lib = load_library()
    arr = (C.c_void_p * len(contexts))(*[c._h for c in contexts])
    if lib.kmn_exchange_reduce_all(arr, len(contexts)) != 0:
        raise KaminoError(lib.kmn_last_error().decode("utf-8", "replace"))


class PinnedBuffer:
    """Page-locked host bytes (kmn_host_alloc) as a numpy array: .array, released by .free()."""

    def __init__(self, nbytes: int):
        self.lib = load_library()
        self._p = C.c_void_p()
        if self.lib.kmn_host_alloc(C.byref(self._p), max(int(nbytes), 1)) != 0:
            raise KaminoError(self.lib.kmn_last_error().decode())
        self.array = np.ctypeslib.as_array(C.cast(self._p, C.POINTER(C.c_uint8)), shape=(int(nbytes),))

    def free(self):
        if self._p:
            self.array = None
            self.lib.kmn_host_free(self._p)
            self._p = None

    @classmethod
    def copy_of(cls, a: np.ndarray) -> tuple["PinnedBuffer", np.ndarray]:
        """A page-locked copy of `a`: (buffer to free, view with a's dtype and shape).  cudaMemcpyAsync from pageable
        memory is staged by the driver and blocks the calling thread; offsets handed to kmn_count_batch should be
        pinned like the residues."""
        a = np.ascontiguousarray(a)
        buf = cls(a.nbytes)
        view = buf.array[: a.nbytes].view(a.dtype).reshape(a.shape)
        view[...] = a
        return buf, view


class Context:
    """One kmn_ctx: one GPU, one table (mirrors AtomicCountMinSketch + CountMinSketch ownership)."""

    def __init__(self, k: int = 13, scheme: int | str = SR6, device: int = 0):
        self.lib = load_library()
        if isinstance(scheme, str):
            scheme = SCHEMES[scheme.lower()]
        self.k, self.scheme, self.device = k, scheme, device
        h = C.c_void_p()
        self._check(self.lib.kmn_create(C.byref(h), device, k, scheme))
        self._h = h

    def _check(self, rc: int):
        if rc != 0:
            raise KaminoError(self.lib.kmn_last_error().decode("utf-8", "replace"))

    def close(self):
        if getattr(self, "_h", None):
            self.lib.kmn_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- pass 1
    def stage_batch(self, residues, rec_off, sp_rec_off) -> int:
        residues = _as(residues, np.uint8)
        rec_off, sp_rec_off = _as(rec_off, np.uint64), _as(sp_rec_off, np.uint64)
        bid = C.c_uint32()
        self._check(self.lib.kmn_stage_batch(self._h, _ptr(residues), _ptr(rec_off), len(rec_off) - 1,
                                             _ptr(sp_rec_off), len(sp_rec_off) - 1, C.byref(bid)))
        return bid.value

    def stage_batch_async(self, residues, rec_off, sp_rec_off) -> int:
        """kmn_stage_batch_async: queues the copies and returns.  The arrays must be contiguous with the right dtype
        already (no hidden temporary copy may be handed to an asynchronous copy) and are kept referenced until the
        batch is released; pinned arrays (PinnedBuffer) are needed for the copies to overlap anything."""
        for a, dt in ((residues, np.uint8), (rec_off, np.uint64), (sp_rec_off, np.uint64)):
            if not (isinstance(a, np.ndarray) and a.dtype == dt and a.flags.c_contiguous):
                raise KaminoError("stage_batch_async needs contiguous numpy arrays of uint8 / uint64 / uint64")
        bid = C.c_uint32()
        self._check(self.lib.kmn_stage_batch_async(self._h, _ptr(residues), _ptr(rec_off), len(rec_off) - 1,
                                                   _ptr(sp_rec_off), len(sp_rec_off) - 1, C.byref(bid)))
        if not hasattr(self, "_async_refs"):
            self._async_refs = {}
        self._async_refs[bid.value] = (residues, rec_off, sp_rec_off)
        return bid.value

    def release_batch(self, batch_id: int):
        self._check(self.lib.kmn_release_batch(self._h, batch_id))
        getattr(self, "_async_refs", {}).pop(batch_id, None)

    def count_staged(self, batch_id: int, n_sp: int | None = None) -> np.ndarray | None:
        out = np.zeros(n_sp, dtype=np.uint64) if n_sp is not None else None
        self._check(self.lib.kmn_count_staged(self._h, batch_id, _ptr(out)))
        return out

    def count_batch(self, residues, rec_off, sp_rec_off, want_counts: bool = True):
        residues = _as(residues, np.uint8)
        rec_off, sp_rec_off = _as(rec_off, np.uint64), _as(sp_rec_off, np.uint64)
        n_sp = len(sp_rec_off) - 1
        out = np.zeros(n_sp, dtype=np.uint64) if want_counts else None
        bid = C.c_uint32()
        self._check(self.lib.kmn_count_batch(self._h, _ptr(residues), _ptr(rec_off), len(rec_off) - 1,
                                             _ptr(sp_rec_off), n_sp, _ptr(out), C.byref(bid)))
        return bid.value, out

    def count_fasta(self, data, file_off, want_counts: bool = True):
        """Whole FASTA files back to back (every non-empty file ends with a line feed): record detection on
        the device, then pass 1.  Returns (batch id, new k-mers per species or None, number of records)."""
        data = _as(np.frombuffer(data, np.uint8) if isinstance(data, (bytes, bytearray)) else data, np.uint8)
        file_off = _as(file_off, np.uint64)
        n_files = len(file_off) - 1
        new = np.zeros(n_files, dtype=np.uint64) if want_counts else None
        bid, n_rec = C.c_uint32(), C.c_uint64()
        self._check(self.lib.kmn_count_fasta(self._h, _ptr(data), _ptr(file_off), n_files, _ptr(new), C.byref(bid), C.byref(n_rec)))
        return bid.value, new, int(n_rec.value)

    def fetch_records(self, batch_id: int, n_files: int) -> tuple[np.ndarray, np.ndarray]:
        """(offset of each record's '>' in the bytes given to count_fasta, first record of each file)."""
        n = C.c_uint64()
        sp = np.zeros(n_files + 1, dtype=np.uint64)
        self._check(self.lib.kmn_fetch_records(self._h, batch_id, None, 0, C.byref(n), _ptr(sp)))
        hdr = np.zeros(int(n.value), dtype=np.uint64)
        self._check(self.lib.kmn_fetch_records(self._h, batch_id, _ptr(hdr), n.value, C.byref(n), None))
        return hdr, sp

    def reset_table(self):
        self._check(self.lib.kmn_reset_table(self._h))

    def release_batches(self):
        self._check(self.lib.kmn_release_batches(self._h))
        getattr(self, "_async_refs", {}).clear()

    def finalize_table(self):
        self._check(self.lib.kmn_finalize_table(self._h))

    def table_snapshot(self) -> np.ndarray:
        out = np.empty(CMS_CELLS, dtype=np.uint16)
        self._check(self.lib.kmn_table_snapshot(self._h, _ptr(out)))
        return out

    def table_load(self, table):
        table = _as(table, np.uint16)
        assert table.size == CMS_CELLS
        self._check(self.lib.kmn_table_load(self._h, _ptr(table)))

    def accum_device_ptr(self) -> int:
        p = C.c_void_p()
        self._check(self.lib.kmn_table_accum_device_ptr(self._h, C.byref(p)))
        return p.value

    def table_widen(self):
        self._check(self.lib.kmn_table_widen(self._h))

    def table_device_ptr(self) -> int:
        p = C.c_void_p()
        self._check(self.lib.kmn_table_device_ptr(self._h, C.byref(p)))
        return p.value

    def exchange_export(self) -> tuple[bytes, bytes]:
        a, t = np.zeros(64, np.uint8), np.zeros(64, np.uint8)
        self._check(self.lib.kmn_exchange_export(self._h, _ptr(a), _ptr(t)))
        return a.tobytes(), t.tobytes()

    def exchange_setup(self, world: int, rank: int, acc_handles: list[bytes], table_handles: list[bytes]):
        a = np.frombuffer(b"".join(acc_handles), dtype=np.uint8).copy()
        t = np.frombuffer(b"".join(table_handles), dtype=np.uint8).copy()
        assert a.size == 64 * world and t.size == 64 * world
        self._check(self.lib.kmn_exchange_setup(self._h, world, rank, _ptr(a), _ptr(t)))

    def exchange_reduce(self):
        self._check(self.lib.kmn_exchange_reduce(self._h))

    def exchange_abort(self):
        self._check(self.lib.kmn_exchange_abort(self._h))

    def table_checksum(self) -> tuple[np.ndarray, int]:
        """(row sums [4] u64, position-sensitive 64-bit hash) of the table, reduced on the device."""
        rows, h = np.zeros(4, np.uint64), C.c_uint64()
        self._check(self.lib.kmn_table_checksum(self._h, _ptr(rows), C.byref(h)))
        return rows, int(h.value)

    def stream(self) -> int:
        p = C.c_void_p()
        self._check(self.lib.kmn_stream(self._h, C.byref(p)))
        return p.value or 0

    def synchronize(self):
        self._check(self.lib.kmn_synchronize(self._h))

    # ---- pass 2
    def scan_batch(self, batch_id: int, min_needed: int) -> tuple[int, int]:
        ns, ne = C.c_uint64(), C.c_uint64()
        self._check(self.lib.kmn_scan_batch(self._h, batch_id, min_needed, C.byref(ns), C.byref(ne)))
        return ns.value, ne.value

    def fetch_hits(self, batch_id: int, sp: int) -> tuple[np.ndarray, np.ndarray]:
        ns, ne = C.c_uint64(), C.c_uint64()
        self._check(self.lib.kmn_fetch_hits(self._h, batch_id, sp, None, 0, C.byref(ns), None, 0, C.byref(ne)))
        starts = np.zeros(ns.value, dtype=HIT_DTYPE)
        ends = np.zeros(ne.value, dtype=HIT_DTYPE)
        self._check(self.lib.kmn_fetch_hits(self._h, batch_id, sp, _ptr(starts), ns.value, C.byref(ns),
                                            _ptr(ends), ne.value, C.byref(ne)))
        return starts, ends

    def pair_batch(self, batch_id: int, length_middle: int) -> int:
        n = C.c_uint64()
        self._check(self.lib.kmn_pair_batch(self._h, batch_id, length_middle, C.byref(n)))
        return n.value

    def fetch_pairs(self, batch_id: int, sp: int) -> np.ndarray:
        n = C.c_uint64()
        self._check(self.lib.kmn_fetch_pairs(self._h, batch_id, sp, None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=PAIR_DTYPE)
        if n.value:
            self._check(self.lib.kmn_fetch_pairs(self._h, batch_id, sp, _ptr(out), n.value, C.byref(n)))
        return out

    def group_observations(self, left, right, sid, length, min_needed: int) -> tuple[np.ndarray, np.ndarray]:
        left, right = _as(left, np.uint64), _as(right, np.uint64)
        sid, length = _as(sid, np.uint32), _as(length, np.uint32)
        n = left.size
        perm = np.zeros(n, dtype=np.uint32)
        groups = np.zeros(max(n, 1), dtype=GROUP_DTYPE)
        ng = C.c_uint64()
        self._check(self.lib.kmn_group_observations(self._h, n, _ptr(left), _ptr(right), _ptr(sid), _ptr(length),
                                                    min_needed, _ptr(perm), _ptr(groups), groups.size, C.byref(ng)))
        return perm, groups[:ng.value]

    def group_batches(self, batch_ids, min_needed: int, sid_base=None, sid_stride: int = 1, full: bool = False):
        """Grouping of the anchor pairs of paired batches, read in place on the device (kmn_group_batches).
        Returns (n_obs, n_groups, n_valid), or (perm, groups) with full=True."""
        ids = _as(batch_ids, np.uint32)
        if sid_base is None:
            sid_base = np.zeros(ids.size, np.uint32)
        base = _as(sid_base, np.uint32)
        n_obs, ng = C.c_uint64(), C.c_uint64()
        self._check(self.lib.kmn_group_batches(self._h, _ptr(ids), ids.size, _ptr(base), sid_stride, min_needed, None, 0, None, 0,
                                               C.byref(n_obs), C.byref(ng)))
        n = int(n_obs.value)
        perm = np.zeros(max(n, 1), dtype=np.uint32)
        groups = np.zeros(max(n, 1), dtype=GROUP_DTYPE)
        self._check(self.lib.kmn_group_batches(self._h, _ptr(ids), ids.size, _ptr(base), sid_stride, min_needed, _ptr(perm), perm.size,
                                               _ptr(groups), groups.size, C.byref(n_obs), C.byref(ng)))
        groups = groups[:ng.value]
        if full:
            return perm[:n], groups
        return n, int(ng.value), int(np.count_nonzero(groups["flags"] & GROUP_VALID))

    def fetch_occupancy(self, batch_id: int, sp: int) -> np.ndarray:
        n = C.c_uint64()
        self._check(self.lib.kmn_fetch_occupancy(self._h, batch_id, sp, None, 0, C.byref(n)))
        occ = np.zeros(n.value, dtype=np.uint16)
        self._check(self.lib.kmn_fetch_occupancy(self._h, batch_id, sp, _ptr(occ), n.value, C.byref(n)))
        return occ

    # ---- --nj
    def pair_counts(self, mapped: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
        mapped = _as(mapped, np.uint8)
        n, L = mapped.shape
        valid = np.zeros((n, n), dtype=np.uint32)
        mism = np.zeros((n, n), dtype=np.uint32)
        self._check(self.lib.kmn_pair_counts(self._h, _ptr(mapped), n, L, _ptr(valid), _ptr(mism)))
        return valid, mism

    def pair_counts_rows(self, mapped: np.ndarray, row_begin: int, row_end: int, fetch: bool = True):
        """Rows [row_begin, row_end) of the pair matrix (one rank's block of the sharded --nj counts).
        fetch=False leaves the counts on the device (pair_counts_device) and returns None."""
        mapped = _as(mapped, np.uint8)
        n, L = mapped.shape
        rows = max(row_end - row_begin, 0)
        if not fetch:
            self._check(self.lib.kmn_pair_counts_rows(self._h, _ptr(mapped), n, L, row_begin, row_end, None, None))
            return None
        valid = np.zeros((rows, n), dtype=np.uint32)
        mism = np.zeros((rows, n), dtype=np.uint32)
        self._check(self.lib.kmn_pair_counts_rows(self._h, _ptr(mapped), n, L, row_begin, row_end, _ptr(valid), _ptr(mism)))
        return valid, mism

    def pair_counts_device(self) -> tuple[int, int, int, int]:
        """(valid device pointer, mismatch device pointer, rows, n) of the last pair-count call."""
        v, m, rows, n = C.c_void_p(), C.c_void_p(), C.c_uint32(), C.c_uint32()
        self._check(self.lib.kmn_pair_counts_device(self._h, C.byref(v), C.byref(m), C.byref(rows), C.byref(n)))
        return int(v.value or 0), int(m.value or 0), rows.value, n.value

    # ---- instrumentation
    def stage_ms(self, name: str) -> float:
        ms = C.c_float()
        self._check(self.lib.kmn_last_stage_ms(self._h, STAGES[name], C.byref(ms)))
        return ms.value

    def launch_count(self) -> int:
        return int(self.lib.kmn_launch_count(self._h))
