"""Multi-GPU layer: one process per GPU, contiguous shards, counts gathered over NCCL.

The matching path shards naturally: documents are independent (C3) and a
single huge buffer splits into byte ranges that only need a small halo (C1/C2) or a cut at a
byte no match can straddle (C4/C5).  No text crosses the links; the only exchange is one
all-gather of a small per-rank record {n_units, out_bytes, first_span, last_span} from which
every rank derives the global offset of its slice of the result ("global compaction") and,
for class-run patterns, stitches a run that touches a shard boundary.

`torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is the plumbing; the math on the
gathered records is plain Python so that the world_size-2 gloo tests exercise it.
"""

from __future__ import annotations

import os
import time
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch

HALO = 1 << 16


@dataclass
class Shard:
    config: str
    buffer: torch.Tensor        # uint8 [left halo | core | right halo]
    core_begin: int             # in buffer coordinates
    core_end: int
    global_offset: int          # position of buffer[0] in the whole corpus
    rank: int
    world: int
    n_docs: int = 0
    text: object = None         # mr.Text over `buffer` with the window set (GPU only)

    @property
    def core_bytes(self) -> int:
        return self.core_end - self.core_begin


def shard_bounds(total: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, near-equal byte ranges."""
    return [((total * r) // world, (total * (r + 1)) // world) for r in range(world)]


def _core(config: str, size: int, rank: int, device):
    from . import corpora

    base = {"c1": 0xC1, "c2": 0xC2, "c3": 0xC3, "c4": 0xC4, "c5": 0xC5}[config]
    return corpora.make(config, size, seed=base + 1000003 * rank, device=device)


def make_shard(config: str, size: int, rank: int, world: int, device, with_text: bool = True) -> Shard:
    """Weak-scaling shard: every rank owns `size` units (bytes; lines for c3) of a world*size corpus."""
    core = _core(config, size, rank, device)
    if config == "c3":
        buf, cb, ce, goff, n_docs = core, 0, core.numel(), rank * core.numel(), size
    elif config in ("c4", "c5") or world == 1:
        # sub/split assemble per shard: end every shard on a byte no match can straddle
        if world > 1:
            core[-1] = 32 if config == "c4" else 10
        buf, cb, ce, goff, n_docs = core, 0, core.numel(), rank * core.numel(), 0
    else:
        n = core.numel()
        h = min(HALO, n)
        left = _core(config, size, rank - 1, device)[n - h:] if rank > 0 else core[:0]
        right = _core(config, size, rank + 1, device)[:h] if rank + 1 < world else core[:0]
        buf = torch.cat([left, core, right])
        cb, ce, goff, n_docs = left.numel(), left.numel() + n, rank * n - left.numel(), 0
    sh = Shard(config, buf.contiguous(), cb, ce, goff, rank, world, n_docs)
    if with_text and buf.is_cuda:
        import mojo_regex_b200 as mr

        sh.text = mr.Text.from_tensor(sh.buffer)
        sh.text.set_window(cb, ce, goff)
    return sh


# ---- the exchanged record --------------------------------------------------------------------
def gather_records(rec: List[int], device) -> List[List[int]]:
    """All-gather one small int64 record per rank (the only collective on the path)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [list(rec)]
    t = torch.tensor(rec, dtype=torch.int64, device=device)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [o.tolist() for o in out]


class DeviceGather:
    """All-gather of a small int64 record that never blocks the scan stream -- nor the thread that launches
    the scans.

    The record (match count, output bytes) is known on the host when a matching call returns, so the
    exchange has no dependency on the kernels.  It is issued by a worker thread (FIFO, so every rank issues
    its collectives in the same order) on the gather's own CUDA stream: H2D of the record from a ring of
    pinned slots, one NCCL all-gather, and, on demand, an exclusive scan on the device.  The launching
    thread goes straight on to the next step's scan -- which spends its time inside the C library with the
    GIL released -- so neither the ~0.1 ms of host work per collective nor its latency sits between two
    scans.  Consumers of `base_offset()`, and anything that times the steps, call `join(stream)`."""

    RING = 64

    def __init__(self, width: int, device):
        import queue
        import threading

        import torch.distributed as dist

        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.cuda = torch.device(device).type == "cuda"
        self.device = device
        self.rec = torch.zeros(width, dtype=torch.int64, device=device)
        self.table = torch.zeros((self.world, width), dtype=torch.int64, device=device)
        self.host = torch.zeros((self.RING, width), dtype=torch.int64)
        if self.cuda:
            self.host = self.host.pin_memory()
        self.stream = torch.cuda.Stream(device=device) if self.cuda else None
        self.i = 0
        self._q = queue.Queue()
        self._err = None
        self._worker = threading.Thread(target=self._run, name="mrx-exchange", daemon=True)
        self._worker.start()

    def _issue(self, slot) -> None:
        import contextlib

        import torch.distributed as dist

        ctx = torch.cuda.stream(self.stream) if self.cuda else contextlib.nullcontext()
        with ctx:
            self.rec.copy_(slot, non_blocking=True)
            if self.world > 1:
                dist.all_gather_into_tensor(self.table.view(-1), self.rec)
            else:
                self.table[0].copy_(self.rec)

    def _run(self) -> None:
        if self.cuda:
            torch.cuda.set_device(self.device)
        while True:
            slot = self._q.get()
            try:
                if slot is not None and self._err is None:
                    self._issue(slot)
            except Exception as e:  # surfaced by the next drain()
                self._err = e
            finally:
                self._q.task_done()

    def gather(self, values: List[int]) -> torch.Tensor:
        """Queue one exchange; -> the (world, width) device table (valid on self.stream after drain())."""
        if self.i - self._done_floor() >= self.RING - 1:
            self.drain()  # never overwrite a pinned slot whose copy may still be pending
        slot = self.host[self.i % self.RING]
        self.i += 1
        for k, v in enumerate(values):
            slot[k] = int(v)
        self._q.put(slot)
        return self.table

    def _done_floor(self) -> int:
        return self.i - self._q.unfinished_tasks

    def drain(self) -> None:
        """Wait until every queued exchange has been ISSUED (not necessarily finished on the device)."""
        self._q.join()
        if self._err is not None:
            err, self._err = self._err, None
            raise err

    def base_offset(self, column: int = 0) -> torch.Tensor:
        """Exclusive prefix of `column` at this rank (0-d device tensor, valid on self.stream)."""
        import contextlib

        self.drain()
        ctx = torch.cuda.stream(self.stream) if self.cuda else contextlib.nullcontext()
        with ctx:
            col = self.table[:, column]
            return torch.cumsum(col, 0)[self.rank] - col[self.rank]

    def join(self, stream) -> None:
        """Make `stream` wait for every exchange issued so far."""
        self.drain()
        if self.cuda:
            stream.wait_stream(self.stream)


class NativeGather:
    """The same exchange through the library's own NCCL communicator (mrx_comm_*): a few microseconds of
    host time per step in C, no Python thread, no torch dispatcher.  Used when `init_native_exchange`
    succeeded on every rank; `DeviceGather` (torch.distributed) is the fallback."""

    def __init__(self, width: int, device):
        import torch.distributed as dist

        self.width = width
        self.device = device
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()
        self.table_ptr = 0

    def gather(self, values: List[int]) -> None:
        import mojo_regex_b200 as mr

        assert len(values) == self.width
        self.table_ptr = mr.comm_allgather(values)

    def gather_device(self, value_ptr: int) -> None:
        """The same for one int64 that is still being produced on the library stream (device address)."""
        import mojo_regex_b200 as mr

        assert self.width == 1
        self.table_ptr = mr.comm_allgather_device(value_ptr)

    def table(self) -> torch.Tensor:
        """(world, width) view of the library's table, ordered after the exchanges on the current stream."""
        import mojo_regex_b200 as mr

        mr.comm_join(torch.cuda.current_stream().cuda_stream)

        class _V:
            pass

        v = _V()
        v.__cuda_array_interface__ = {"shape": (self.world, self.width), "typestr": "<i8", "data": (self.table_ptr, False),
                                      "version": 3, "strides": None}
        return torch.as_tensor(v, device=self.device)

    def base_offset(self, column: int = 0) -> torch.Tensor:
        col = self.table()[:, column]
        return torch.cumsum(col, 0)[self.rank] - col[self.rank]

    def drain(self) -> None:
        pass

    def join(self, stream) -> None:
        import mojo_regex_b200 as mr

        mr.comm_join(stream.cuda_stream)


_NATIVE = False


def init_native_exchange(device) -> bool:
    """Set up the library's NCCL communicator next to torch.distributed's (rank 0 creates the id, it is
    broadcast through torch.distributed).  All ranks agree on the outcome; False = keep DeviceGather."""
    global _NATIVE
    import torch.distributed as dist

    import mojo_regex_b200 as mr

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return False
    # probe first (not collective): ncclCommInitRank below blocks until EVERY rank has called it, so nobody
    # may enter it unless every rank can
    try:
        probe_id = mr.comm_unique_id()
        probe = 1
    except Exception:
        probe_id, probe = b"", 0
    pf = torch.tensor([probe], dtype=torch.int32, device=device)
    dist.all_reduce(pf, op=dist.ReduceOp.MIN)
    if not int(pf.item()):
        return False
    ok = 1
    try:
        rank, world = dist.get_rank(), dist.get_world_size()
        uid = torch.zeros(128, dtype=torch.uint8, device=device)
        if rank == 0:
            uid.copy_(torch.tensor(list(probe_id), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        if ok:
            try:
                if int(uid.sum().item()) == 0:
                    ok = 0
                else:
                    mr.comm_init(bytes(uid.cpu().tolist()), rank, world)
            except Exception:
                ok = 0
    except Exception:
        ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    _NATIVE = bool(flag.item())
    return _NATIVE


def global_offsets(counts: List[int]) -> List[int]:
    """Exclusive prefix sum: where each rank's slice of the global result starts."""
    out, acc = [], 0
    for c in counts:
        out.append(acc)
        acc += c
    return out


def stitch_class_runs(records: List[List[int]]) -> Tuple[List[int], List[Optional[Tuple[int, int]]], List[bool]]:
    """CLASS_RUNS shards report runs clipped to their core.  Record = [n, first_s, first_e, last_s,
    last_e, core_begin_global, core_end_global].  A run that ends exactly on a shard's end and a run
    that starts exactly on the next shard's begin are one run of the reference.

    Returns (counts after stitching, replacement for each rank's last span or None, drop_first flags).
    """
    world = len(records)
    counts = [r[0] for r in records]
    new_last: List[Optional[Tuple[int, int]]] = [None] * world
    drop_first = [False] * world
    r = 0
    while r < world:
        n, fs, fe, ls, le, cb, ce = records[r]
        if n == 0 or le != ce:
            r += 1
            continue
        # run open at the end of rank r: extend through following ranks
        end, k = le, r + 1
        while k < world:
            nk, fsk, fek, lsk, lek, cbk, cek = records[k]
            if nk == 0 or fsk != cbk or fsk != end:
                break
            drop_first[k] = True
            counts[k] -= 1
            end = fek
            if not (nk == 1 and fek == cek):
                break
            k += 1
        if end != le:
            new_last[r] = (ls, end)
        # a rank whose single run was swallowed and continues is skipped by the loop above
        r = max(k, r + 1) if end != le else r + 1
    return counts, new_last, drop_first


def findall_record(spans, shard: Shard) -> List[int]:
    """The per-rank record exchanged for findall: [n, first_s, first_e, last_s, last_e, core_begin_g, core_end_g]
    (global coordinates; spans is an (n, 2) array-like of this rank's matches)."""
    n = int(len(spans))
    cb, ce = shard.global_offset + shard.core_begin, shard.global_offset + shard.core_end
    if n == 0:
        return [0, -1, -1, -1, -1, cb, ce]
    return [n, int(spans[0][0]), int(spans[0][1]), int(spans[-1][0]), int(spans[-1][1]), cb, ce]


def finish_findall(spans, records: List[List[int]], rank: int, class_runs: bool, direct: bool = True):
    """Turn a rank's local span list into its slice of the global result.

    Per-shard key-window results compose into the whole-text result only when no match decision crosses a
    shard edge: `direct` plans (mrx_info.findall_direct: potential matches provably never overlap) and the
    stitched CLASS_RUNS case.  Greedy-restart plans are refused here rather than silently mis-sharded.

    Returns (spans_for_this_rank, global_index_of_first_span, total_matches).  For class-run patterns
    runs were clipped at the shard edges and are stitched here; for every other mode shards already
    report exactly the matches whose key lies in their core, so only offsets are needed."""
    import numpy as np

    if not (class_runs or direct):
        raise ValueError("sharded findall needs a direct plan (mrx_info.findall_direct) or CLASS_RUNS: a greedy "
                         "restart chain cannot be cut at a shard boundary")
    spans = np.asarray(spans, dtype=np.int64).reshape(-1, 2)
    if class_runs:
        counts, new_last, drop_first = stitch_class_runs(records)
        if drop_first[rank] and len(spans):
            spans = spans[1:]
        if new_last[rank] is not None and len(spans):
            spans = spans.copy()
            spans[-1] = new_last[rank]
    else:
        counts = [r[0] for r in records]
    offs = global_offsets(counts)
    return spans, offs[rank], sum(counts)


# ---- bench helpers (GPU) ---------------------------------------------------------------------------
_GATHERS = {}


def join_exchanges(stream) -> None:
    """Order `stream` after every count exchange issued so far (end of a timed region, or a consumer)."""
    for g in _GATHERS.values():
        g.join(stream)


def _gather_for(width: int, device) -> DeviceGather:
    key = (width, str(device))
    if key not in _GATHERS:
        _GATHERS[key] = NativeGather(width, device) if _NATIVE else DeviceGather(width, device)
    return _GATHERS[key]


class OpResult:
    def __init__(self, units: int, algorithmic_bytes: int, kernel: str, holder=None, gather=None):
        self.units = units
        self.algorithmic_bytes = algorithmic_bytes
        self.kernel = kernel
        self.holder = holder
        self._gather = gather

    @property
    def global_base(self):
        """Global index / byte offset of this rank's slice of the result (0-d device tensor on the exchange
        stream, computed on demand from the gathered counts)."""
        return self._gather.base_offset() if self._gather is not None else 0

    def wait(self) -> "OpResult":
        """An asynchronous findall step: fetch the span count (one synchronisation) and fill in units / bytes."""
        if self.units < 0 and self.holder is not None:
            if not self.holder.wait():
                raise RuntimeError("asynchronous findall: the span array was too small; repeat the step")
            self.units = self.holder.n
            self.algorithmic_bytes = self._n_text + 16 * self.holder.n
        return self

    def close(self):
        if self.holder is not None and hasattr(self.holder, "close"):
            self.holder.close()
        self.holder = None


def run_op_local(rx, cfg, shard: Shard) -> OpResult:
    """run_op_device without the cross-rank exchange (safe to call a rank-dependent number of times)."""
    return run_op_device(rx, cfg, shard, exchange=False)


def run_op_device(rx, cfg, shard: Shard, exchange: bool = True, asynchronous: bool = False) -> OpResult:
    """One step with the shard resident in HBM and the result left in HBM.  asynchronous=True (findall): the call
    returns with the scan queued; OpResult.wait() fetches the count."""
    n = shard.core_bytes
    dev = shard.buffer.device
    multi = exchange and shard.world > 1
    if cfg["op"] == "findall":
        g = None
        if asynchronous and (not multi or _NATIVE):
            # no host round trip: the count stays on the device and is exchanged from there; OpResult.units and
            # .algorithmic_bytes are filled in by OpResult.wait()
            ds = rx.match_all_async(shard.text)
            if multi:
                g = _gather_for(1, dev)
                g.gather_device(ds.count_ptr)
        else:
            ds = rx.match_all_device(shard.text)
            if multi:
                g = _gather_for(1, dev)
                g.gather([ds.n])
        kernel = getattr(shard, "_findall_kernel", None)
        if kernel is None:
            info = rx.info()
            kernel = ("k_sand" if info.findall_sand and not os.environ.get("MRX_NO_SAND") else
                      "k_dense" if info.findall_dense else ("k_runs" if info.findall_mode in (4, 8) else "k_scan"))
            shard._findall_kernel = kernel
        r = OpResult(ds.n, n + 16 * ds.n, kernel, ds, g)
        r._n_text = n
        return r
    if cfg["op"] == "search_batch":
        # the caller-provided result array lives with the shard (allocated once, like the text)
        out = getattr(shard, "_batch_out", None)
        if out is None:
            out = shard._batch_out = torch.empty((shard.n_docs, 2), dtype=torch.int64, device=dev)
            info = rx.info()
            shard._batch_kernel = ("k_scan" if info.search_rev_states <= 0 else
                                   "k_lines_req" if info.search_skip_byte >= 0 and os.environ.get("MRX_LINES_FILTER", "1")[:1] != "0"
                                   else "k_lines")
        rx.search_batch_device(shard.text, out.data_ptr(), shard.n_docs, doc_stride=128, sync=False)
        if multi:
            _gather_for(1, dev).gather([shard.n_docs])
        return OpResult(shard.n_docs, n + 16 * shard.n_docs, shard._batch_kernel, None)
    if cfg["op"] == "sub":
        ob = rx.sub_device(cfg["repl"], shard.text)
        g = None
        if multi:
            g = _gather_for(1, dev)
            g.gather([ob.n])
        fused = rx.info().sub_fused_width > 0 and not os.environ.get("MRX_NO_FUSED_SUB")
        return OpResult(n, n + ob.n, "k_subst" if fused else "k_runs+k_seg_lengths+k_assemble", ob, g)
    if cfg["op"] == "split":
        pc = rx.split_device(shard.text)
        total = int(pc.offsets_tensor()[-1].item())
        g = None
        if multi:
            g = _gather_for(2, dev)
            g.gather([pc.n, total])
        fused = rx.info().split_fused and not os.environ.get("MRX_NO_FUSED_SPLIT")
        return OpResult(pc.n, n + total + 8 * (pc.n + 1), "k_split" if fused else "k_scan+k_seg_lengths+k_assemble", pc, g)
    raise ValueError(cfg["op"])


_PINNED_OUT = {}


def _pinned_i64(n: int):
    """Page-locked int64 result buffer for batch search (allocated once per size)."""
    if n not in _PINNED_OUT:
        _PINNED_OUT[n] = torch.empty(n, dtype=torch.int64, pin_memory=True)
    return _PINNED_OUT[n].numpy()


def run_op_host(rx, cfg, shard: Shard, host_pinned: torch.Tensor) -> int:
    """One step through the public API from a pinned host buffer; returns bytes copied back."""
    import mojo_regex_b200 as mr

    if cfg["op"] == "findall":
        # one C-ABI call on the host buffer: upload, scan and span download are pipelined inside
        ml = rx.match_all_host(host_pinned.data_ptr(), host_pinned.numel(), shard.core_begin, shard.core_end,
                               shard.global_offset)
        return int(ml.spans.nbytes)
    if cfg["op"] == "sub" and shard.core_begin == 0 and shard.core_end == host_pinned.numel():
        hb = rx.sub_host(cfg["repl"], host_pinned.data_ptr(), host_pinned.numel())
        nb = hb.n
        hb.close()
        return nb
    if cfg["op"] == "split" and shard.core_begin == 0 and shard.core_end == host_pinned.numel():
        hp = rx.split_host(host_pinned.data_ptr(), host_pinned.numel())
        nb = int(hp.data.nbytes) + int(hp.offsets.nbytes)
        hp.close()
        return nb
    if cfg["op"] == "search_batch":
        rx.search_batch_host(host_pinned.data_ptr(), host_pinned.numel(), _pinned_i64(2 * shard.n_docs), doc_stride=128,
                             n_docs=shard.n_docs)
        return int(2 * shard.n_docs * 8)
    t = mr.Text.from_pinned(host_pinned.data_ptr(), host_pinned.numel())
    t.set_window(shard.core_begin, shard.core_end, shard.global_offset)
    try:
        if cfg["op"] == "findall":
            ml = rx.match_all(t)
            return int(ml.spans.nbytes)
        if cfg["op"] == "search_batch":
            out = rx.search_batch(t, doc_stride=128, n_docs=shard.n_docs, out=_pinned_i64(2 * shard.n_docs))
            return int(2 * shard.n_docs * 8)
        if cfg["op"] == "sub":
            hb = rx.sub_view(cfg["repl"], t)
            nb = hb.n
            hb.close()
            return nb
        if cfg["op"] == "split":
            hp = rx.split_view(t)
            nb = int(hp.data.nbytes) + int(hp.offsets.nbytes)
            hp.close()
            return nb
    finally:
        t.close()
    raise ValueError(cfg["op"])
