"""CPU tests of the N>1 path: world_size-2 gloo processes shard the species, build their partial
tables with the oracle, exchange them like the GPU path does (sum of u32 accumulators, then clamp)
and must reproduce the single-process table."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    import oracle_ffi
    from kamino_b200 import sharding, synth

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc = oracle_ffi.load()
        species = synth.bacterial(6, seed=77, n_prot=60, mean_len=120.0)
        b = synth.stage(species)
        mine = sharding.shard_species(b.n_sp, world, rank)
        res, rec, sp = sharding.select_species(b.residues, b.rec_off, b.sp_rec_off, mine)
        h = orc.cmsa_new()
        try:
            new = orc.count_batch(h, res, rec, sp, 13, 1, threads=2)
            table = orc.cmsa_table(h)
            nz = np.flatnonzero(table)
            # sparse exchange keeps the CPU test light: all-gather (index, value) pairs and sum
            acc = torch.zeros(4 << 26, dtype=torch.int32)
            acc[torch.from_numpy(nz)] = torch.from_numpy(table[nz].astype(np.int32))
        finally:
            orc.cmsa_free(h)
        sharding.exchange_accumulators(acc)
        merged = sharding.clamp_u16(acc.numpy())
        counts = [None] * world
        dist.all_gather_object(counts, (mine.tolist(), new.tolist()))
        if rank == 0:
            h = orc.cmsa_new()
            try:
                want_new = orc.count_batch(h, b.residues, b.rec_off, b.sp_rec_off, 13, 1, threads=2)
                want = np.array(orc.cmsa_table(h))
            finally:
                orc.cmsa_free(h)
            got_new = np.zeros(b.n_sp, dtype=np.uint64)
            for ids, vals in counts:
                got_new[ids] = vals
            ok = bool(np.array_equal(merged, want)) and got_new.tolist() == want_new.tolist()
            Path(out_dir, "result.txt").write_text("ok" if ok else "mismatch")
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_round_robin_shards_cover_everything():
    from kamino_b200 import sharding
    for n, w in [(0, 2), (1, 2), (7, 2), (100, 8), (1000, 8), (5, 8)]:
        seen = np.concatenate([sharding.shard_species(n, w, r) for r in range(w)])
        assert sorted(seen.tolist()) == list(range(n))
        sizes = [len(sharding.shard_species(n, w, r)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1


def test_select_species_roundtrip():
    from kamino_b200 import sharding, synth
    b = synth.stage(synth.bacterial(5, seed=3, n_prot=20, mean_len=80.0))
    res, rec, sp = sharding.select_species(b.residues, b.rec_off, b.sp_rec_off, np.arange(b.n_sp))
    assert np.array_equal(res, b.residues) and np.array_equal(rec, b.rec_off) and np.array_equal(sp, b.sp_rec_off)
    res, rec, sp = sharding.select_species(b.residues, b.rec_off, b.sp_rec_off, np.array([3, 1]))
    r0, r1 = int(b.sp_rec_off[3]), int(b.sp_rec_off[4])
    assert np.array_equal(res[:int(rec[int(sp[1])])], b.residues[int(b.rec_off[r0]):int(b.rec_off[r1])])


def test_two_rank_gloo_exchange_matches_single_process(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "result.txt").read_text() == "ok"


def test_clamp_is_exact_under_saturation():
    from kamino_b200 import sharding
    a = np.array([0, 1, 65534, 65535, 65536, 70000, 2**31 - 1], dtype=np.int64).astype(np.int32)
    assert sharding.clamp_u16(a).tolist() == [0, 1, 65534, 65535, 65535, 65535, 65535]
    # min(sum(min(a_r, M)), M) == min(sum(a_r), M): partial clamping commutes with the exchange
    rng = np.random.default_rng(0)
    parts = rng.integers(0, 50000, size=(4, 1000))
    assert np.array_equal(np.minimum(np.minimum(parts, 65535).sum(0), 65535), np.minimum(parts.sum(0), 65535))


def test_pair_row_blocks_cover_the_matrix():
    from kamino_b200 import sharding
    for n, w in [(0, 2), (1, 3), (64, 2), (100, 4), (130, 3), (5000, 1), (5000, 2), (5000, 8), (777, 5)]:
        blocks = sharding.pair_row_blocks(n, w)
        assert len(blocks) == w and blocks[0][0] == 0 and blocks[-1][1] == n
        for (b0, b1), (c0, _) in zip(blocks, blocks[1:]):
            assert b1 == c0                                            # contiguous, disjoint
        for b0, b1 in blocks:
            assert b0 <= b1 and b0 % 64 == 0 and (b1 % 64 == 0 or b1 == n)
    # balanced by upper-triangle tiles, not by rows
    nt = (5000 + 63) // 64
    work = [sum(nt - t for t in range(b0 // 64, (b1 + 63) // 64)) for b0, b1 in sharding.pair_row_blocks(5000, 8)]
    assert max(work) <= 1.12 * (sum(work) / 8)


def _pair_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    import oracle_ffi
    from kamino_b200 import sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc = oracle_ffi.load()
        rng = np.random.default_rng(5)
        n, L = 150, 333
        m = rng.integers(0, 21, size=(n, L), dtype=np.uint8)
        want_v, want_m = orc.pair_counts(m, threads=2)
        iu = np.triu_indices(n, 1)

        def count_rows(mapped, b0, b1):   # the oracle stands in for Context.pair_counts_rows on CPU
            v = np.zeros((b1 - b0, n), np.uint32)
            mm = np.zeros((b1 - b0, n), np.uint32)
            for i in range(b0, b1):
                v[i - b0, i + 1:] = want_v[i, i + 1:]
                mm[i - b0, i + 1:] = want_m[i, i + 1:]
            return v, mm

        got = sharding.pair_counts_sharded(count_rows, m)
        if rank == 0:
            ok = np.array_equal(got[0][iu], want_v[iu]) and np.array_equal(got[1][iu], want_m[iu])
            Path(out_dir, "pair.txt").write_text("ok" if ok else "mismatch")
        else:
            assert got is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_pair_counts_gather(tmp_path):
    mp.spawn(_pair_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "pair.txt").read_text() == "ok"


def test_row_blocks_and_species_shards_properties():
    """Randomised invariants of the two partitioners (hypothesis): exact cover, alignment, balance."""
    from hypothesis import given, settings, strategies as st
    from kamino_b200 import sharding

    @settings(max_examples=300, deadline=None)
    @given(st.integers(0, 70_000), st.integers(1, 16))
    def rows(n, world):
        blocks = sharding.pair_row_blocks(n, world)
        assert len(blocks) == world and blocks[0][0] == 0 and blocks[-1][1] == n
        covered = 0
        for b0, b1 in blocks:
            assert b0 == covered and b1 >= b0 and b0 % 64 == 0 and (b1 % 64 == 0 or b1 == n)
            covered = b1
        nt = (n + 63) // 64
        work = [sum(nt - t for t in range(b0 // 64, (b1 + 63) // 64)) for b0, b1 in blocks]
        assert sum(work) == nt * (nt + 1) // 2
        if nt >= 8 * world:                      # enough row tiles to balance: no rank above 1.5 x the mean
            assert max(work) <= 1.5 * sum(work) / world

    @settings(max_examples=200, deadline=None)
    @given(st.integers(0, 5000), st.integers(1, 16))
    def species(n, world):
        parts = [sharding.shard_species(n, world, r) for r in range(world)]
        assert sorted(np.concatenate(parts).tolist()) == list(range(n))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1

    rows()
    species()
