"""World-size-2 (and emulated world-size-4) checks of the multi-GPU layer on CPU with gloo.

The CUDA path cannot run here, so each rank matches its shard with the oracle restricted to the
shard's key window -- exactly the contract mrx_text_set_window gives the kernels -- and the test
checks that shard construction (halos, cuts), the all-gathered records, the offset scan and the
class-run stitching reproduce the oracle's result on the whole corpus."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from mojo_regex_b200 import corpora
from mojo_regex_b200 import distributed as mrd


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _windowed_oracle(rx, shard, class_runs):
    """What the GPU library returns for a windowed text, computed with the oracle."""
    buf = shard.buffer.numpy()
    if class_runs:  # runs clipped to the core (k_runs treats the key window as the whole text)
        core = buf[shard.core_begin:shard.core_end]
        sp = rx.findall(core) + shard.core_begin
    else:
        sp = rx.findall(buf)
        sp = sp[(sp[:, 0] >= shard.core_begin) & (sp[:, 0] < shard.core_end)]
    return sp + shard.global_offset


def _whole_corpus(config, size, world):
    return torch.cat([mrd._core(config, size, r, "cpu") for r in range(world)]).numpy()


def _worker(rank, world, port, config, size, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pat = corpora.CONFIGS[config]["pattern"]
        rx = oracle.compile(pat)
        class_runs = config == "c1"
        shard = mrd.make_shard(config, size, rank, world, "cpu", with_text=False)
        local = _windowed_oracle(rx, shard, class_runs)
        records = mrd.gather_records(mrd.findall_record(local, shard), "cpu")
        mine, first, total = mrd.finish_findall(local, records, rank, class_runs)
        np.save(os.path.join(out_dir, f"spans_{rank}.npy"), mine)
        np.save(os.path.join(out_dir, f"meta_{rank}.npy"), np.array([first, total]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("config,size", [("c2", 300_000), ("c1", 200_000)])
def test_two_rank_gloo_findall_equals_whole_corpus(tmp_path, config, size):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, config, size, str(tmp_path)), nprocs=world, join=True)
    whole = _whole_corpus(config, size, world)
    want = oracle.compile(corpora.CONFIGS[config]["pattern"]).findall(whole)
    got, total = [], None
    for r in range(world):
        sp = np.load(tmp_path / f"spans_{r}.npy")
        first, total = np.load(tmp_path / f"meta_{r}.npy")
        assert first == sum(len(g) for g in got)  # global offset = number of spans on lower ranks
        got.append(sp)
    got = np.concatenate(got)
    assert total == len(want)
    assert np.array_equal(got, want)


def test_class_run_stitching_across_whole_shards():
    # a digit run that starts in rank 0, swallows all of rank 1 and ends in rank 2
    rx = oracle.compile(r"\d+")
    parts = [b"ab 12 cd 777", b"5555555555", b"999 x 42 y", b"zz 1"]
    whole = np.frombuffer(b"".join(parts), dtype=np.uint8)
    want = rx.findall(whole)
    world, off, records, locals_ = len(parts), 0, [], []
    for r, p in enumerate(parts):
        sh = mrd.Shard("c1", torch.frombuffer(bytearray(p), dtype=torch.uint8), 0, len(p), off, r, world)
        sp = rx.findall(np.frombuffer(p, dtype=np.uint8)) + off
        locals_.append(sp)
        records.append(mrd.findall_record(sp, sh))
        off += len(p)
    got = np.concatenate([mrd.finish_findall(locals_[r], records, r, True)[0] for r in range(world)])
    assert np.array_equal(got, want)
    assert mrd.finish_findall(locals_[0], records, 0, True)[2] == len(want)


def test_shard_bounds_and_offsets():
    assert mrd.shard_bounds(10, 3) == [(0, 3), (3, 6), (6, 10)]
    assert mrd.global_offsets([3, 0, 5]) == [0, 3, 3]
    sh = mrd.make_shard("c2", 100_000, 1, 4, "cpu", with_text=False)
    assert sh.core_bytes == 100_000 and sh.core_begin == min(mrd.HALO, 100_000)
    assert sh.global_offset == 100_000 - sh.core_begin
    # halos are the neighbours' bytes
    left = mrd._core("c2", 100_000, 0, "cpu")
    assert torch.equal(sh.buffer[:sh.core_begin], left[-sh.core_begin:])


def _gather_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = mrd.DeviceGather(2, "cpu")
        bases = []
        for step in range(200):  # more exchanges than pinned slots: the ring must not be overrun
            g.gather([10 * step + rank + 1, 1000 + step])
            if step % 50 == 49:
                bases.append(int(g.base_offset(0)))
        g.join(None)
        np.save(os.path.join(out_dir, f"table_{rank}.npy"), g.table.numpy().copy())
        np.save(os.path.join(out_dir, f"bases_{rank}.npy"), np.array(bases))
    finally:
        dist.destroy_process_group()


def test_device_gather_worker_thread_two_rank_gloo(tmp_path):
    """The count exchange is issued by a worker thread (FIFO): after 200 of them both ranks hold the same
    table, and the exclusive offsets seen along the way are those of the step they were read at."""
    world = 2
    mp.spawn(_gather_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    t0, t1 = np.load(tmp_path / "table_0.npy"), np.load(tmp_path / "table_1.npy")
    assert np.array_equal(t0, t1)
    assert t0.tolist() == [[10 * 199 + 1, 1199], [10 * 199 + 2, 1199]]
    b0, b1 = np.load(tmp_path / "bases_0.npy"), np.load(tmp_path / "bases_1.npy")
    assert b0.tolist() == [0, 0, 0, 0]
    assert b1.tolist() == [10 * s + 1 for s in (49, 99, 149, 199)]
