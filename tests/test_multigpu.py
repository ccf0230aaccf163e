"""Multi-GPU parity at 2 / 4 / 8 ranks (each case skipped unless that many CUDA devices are visible): species sharded over two ranks, tables merged by
the fused NVLink exchange kernel and by the NCCL all-reduce path; both must equal the single-process
oracle table bit for bit."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _DevArray:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def _worker(rank, world, port, out_dir, planes):
    import torch
    import torch.distributed as dist
    import oracle_ffi
    from kamino_b200 import ffi, sharding, synth

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["KMN_EXCHANGE_PLANES"] = planes   # "1": byte-plane exchange (default), "0": whole u16 cells
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        b = synth.stage(synth.bacterial(7, seed=99, n_prot=200, mean_len=150.0))
        mine = sharding.shard_species(b.n_sp, world, rank)
        res, rec, sp = sharding.select_species(b.residues, b.rec_off, b.sp_rec_off, mine)
        results = {}
        with ffi.Context(13, "sr6", device=rank) as ctx:
            fused_ok = sharding.setup_fused_exchange(ctx)
            for mode in (["fused"] if fused_ok else []) + ["nccl"]:
                ctx.reset_table()
                ctx.count_batch(res, rec, sp)
                if mode == "fused":
                    sharding.fused_exchange(ctx)
                else:
                    acc = torch.as_tensor(_DevArray(ctx.accum_device_ptr(), ffi.CMS_CELLS, "<i4"), device=f"cuda:{rank}")
                    ctx.table_widen()                      # u16 cells -> summable u32 copy
                    sharding.exchange_accumulators(acc)
                    torch.cuda.synchronize()
                    ctx.finalize_table()
                results[mode] = ctx.table_snapshot()
                ctx.release_batches()
                dist.barrier()
            # large cells: every rank holds 35 000 species made of one and the same 13-mer, so its four cells
            # are 35 000 per rank (hi bytes in play) and 70 000 -> 65 535 after the exchange (saturation)
            big_ok = True
            if fused_ok:
                one = np.frombuffer(b"ACDEFGHIKLMNP", np.uint8)
                n_big = 35_000
                ctx.reset_table()
                ctx.count_batch(np.tile(one, n_big), np.arange(n_big + 1, dtype=np.uint64) * 13,
                                np.arange(n_big + 1, dtype=np.uint64), want_counts=False)
                ctx.count_batch(res, rec, sp)
                sharding.fused_exchange(ctx)
                t = ctx.table_snapshot()
                orc0 = oracle_ffi.load()
                idx = orc0.cms_indices(int(orc0.roll_kmers(one.tobytes(), 13, 1)[0]))
                cells = [r * (1 << 26) + int(idx[r]) for r in range(4)]
                base = results["fused"].astype(np.int64)
                want_big = base.copy()
                want_big[cells] = np.minimum(base[cells] + world * n_big, 65535)
                big_ok = bool(np.array_equal(t.astype(np.int64), want_big))
                ctx.release_batches()
                dist.barrier()
        orc = oracle_ffi.load()
        h = orc.cmsa_new()
        try:
            orc.count_batch(h, b.residues, b.rec_off, b.sp_rec_off, 13, 1, threads=4)
            want = orc.cmsa_table(h)
            ok = all(np.array_equal(t, want) for t in results.values()) and big_ok
        finally:
            orc.cmsa_free(h)
        Path(out_dir, f"rank{rank}.txt").write_text(("ok " if ok else "mismatch ") + ",".join(results))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("planes", ["1", "0"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_exchange_matches_oracle(tmp_path, world, planes):
    """exchange_planes_kernel<2/4/8> and exchange_reduce_kernel<2,4>/<4,2>/<8,2> against the single-process oracle."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), planes), nprocs=world, join=True)
    for r in range(world):
        txt = (tmp_path / f"rank{r}.txt").read_text()
        assert txt.startswith("ok"), txt
        print("rank", r, txt)


def _timeout_worker(rank, world, port, out_dir):
    import time
    import torch
    import torch.distributed as dist
    from kamino_b200 import ffi, sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["KMN_EXCHANGE_TIMEOUT_MS"] = "1500"
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        with ffi.Context(13, "sr6", device=rank) as ctx:
            assert sharding.setup_fused_exchange(ctx)
            msg = "not run"
            if rank == 0:   # rank 1 never shows up: rank 0 must come back with an error, not hang
                t0 = time.time()
                try:
                    ctx.exchange_reduce()
                    msg = "no error"
                except ffi.KaminoError as e:
                    msg = f"{time.time() - t0:.1f}s {e}"
            dist.barrier()
            # the host can also break the wait at once (kmn_exchange_abort from another thread)
            amsg = "not run"
            assert sharding.setup_fused_exchange(ctx)
            if rank == 0:
                import threading
                box = {}

                def blocked():
                    t0 = time.time()
                    try:
                        ctx.exchange_reduce()
                        box["msg"] = "no error"
                    except ffi.KaminoError as e:
                        box["msg"] = f"{time.time() - t0:.2f}s {e}"

                th = threading.Thread(target=blocked)
                th.start()
                time.sleep(0.3)
                ctx.exchange_abort()
                th.join()
                amsg = box["msg"]
            dist.barrier()
            Path(out_dir, f"a{rank}.txt").write_text(amsg)
            # both ranks set up again and a normal exchange works
            assert sharding.setup_fused_exchange(ctx)
            ctx.reset_table()
            ctx.exchange_reduce()
            rows, _ = ctx.table_checksum()
            Path(out_dir, f"t{rank}.txt").write_text(f"{msg}|{int(rows.sum())}")
            dist.barrier()
    finally:
        dist.destroy_process_group()


def test_exchange_timeout_names_the_missing_rank(tmp_path):
    """A peer that never launches its exchange kernel must not hang the other GPUs (bounded in-kernel waits)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_timeout_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    msg, total = (tmp_path / "t0.txt").read_text().split("|")
    assert "timed out" in msg and "rank 1" in msg, msg
    amsg = (tmp_path / "a0.txt").read_text()
    assert "aborted by the host" in amsg and float(amsg.split("s ")[0]) < 1.2, amsg
    assert total == "0" and (tmp_path / "t1.txt").read_text().endswith("|0")


def _pair_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    import oracle_ffi
    from kamino_b200 import ffi, sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        rng = np.random.default_rng(17)
        n, L = 300, 1999
        m = rng.integers(0, 21, size=(n, L), dtype=np.uint8)
        with ffi.Context(13, "sr6", device=rank) as ctx:
            got = sharding.pair_counts_sharded(ctx, m)
        if rank == 0:
            wv, wm = oracle_ffi.load().pair_counts(m, threads=4)
            iu = np.triu_indices(n, 1)
            ok = np.array_equal(got[0][iu], wv[iu]) and np.array_equal(got[1][iu], wm[iu])
            Path(out_dir, "pair.txt").write_text("ok" if ok else "mismatch")
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_gpu_pair_counts_row_blocks(tmp_path):
    """--nj counts with the pair matrix sharded by row blocks over two GPUs, gathered on rank 0."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_pair_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "pair.txt").read_text() == "ok"
