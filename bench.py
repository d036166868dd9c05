#!/usr/bin/env python3
"""bench.py -- headline benchmark of the matching hot path.

    python bench.py --gpus N --steps K --warmup W [--config c2] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic input.  The headline workload
is BASELINE.json configs[1]: findall("(hello|world|foo)") over a 1 GiB single-buffer corpus.
With N>1 (torchrun, one rank per GPU) every rank scans its own 1 GiB shard (+ halo) of an
N GiB corpus and the ranks all-gather their match counts over NCCL (weak scaling).  The other
four BASELINE configs are measured in the same run and reported under "configs" of the same
JSON line (--configs c2 restricts the run to the headline).

Prints ONE JSON line (rank 0).  `value` = input GB/s with the corpus already resident in HBM
(device time, CUDA events on the library stream, max over ranks); `roofline` = algorithmic bytes of a
step over the dominant kernel's event time (for sub / split: the whole kernel sequence), against the
measured HBM peak, with the DRAM traffic of the committed ncu capture; `e2e` = the same metric through
the host-buffer calls of the C ABI (mrx_findall_host, mrx_search_batch_host, mrx_sub_host, mrx_split_host)
from pinned host memory, every H2D and D2H byte inside the timed region; `cpu_baseline` = the oracle port
of the reference's path on one host thread.  `--impl reference` times that port on all host threads
instead (the reference itself is Mojo and cannot be built here).  With N > 1 the count all-gather runs on
the library's own NCCL communicator (mrx_comm_*), off the scan stream.
"""
import argparse
import collections
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GiB = 1 << 30


KERNEL_FILES = {"k_sand": ["k_sand.cu", "sand.hpp"], "k_split": ["k_split.cu", "sepbits.hpp"], "k_subst": ["k_subst.cu"]}


def kernel_source_hash(kernel: str = "") -> str:
    """sha256 (16 hex digits) over the device code a kernel is built from (its own file and the common headers; the
    kernels without a file of their own live in kernels.cu): an ncu capture is only quoted for the code it was taken on."""
    import hashlib

    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "mojo_regex_b200", "csrc")
    for name in sorted(KERNEL_FILES.get(kernel, ["kernels.cu"]) + ["device.cuh", "kernels.cuh"]):
        with open(os.path.join(csrc, name), "rb") as f:
            h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()[:16]


def load_traffic(config: str, n_bytes: int, kernel: str):
    """DRAM bytes per launch of the dominant kernel (dram__bytes_read.sum + dram__bytes_write.sum of one
    `ncu --set full` capture), from profiles/r02_traffic.json -- quoted only when the capture is of the same
    workload, the same kernel AND the same kernel sources (src_sha); None otherwise (never a stale constant)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            t = json.load(f).get(config)
        if (t and int(t["bytes_per_gpu"]) == int(n_bytes) and t["kernel"] == kernel
                and t.get("src_sha") == kernel_source_hash(kernel)):
            return float(t["traffic_bytes"])
    except Exception:
        pass
    return None


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(prefix="clocks_", suffix=".csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self, t_begin=None, t_end=None):
        """Median SM clock over the samples taken inside [t_begin, t_end] (wall clock); when the timed
        region is shorter than the sampling period, the samples closest to it are used."""
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            import datetime

            rows = [r.strip().split(",") for r in open(self.path) if r.strip()]
            rows = [r for r in rows if len(r) >= 9]
            if t_begin is not None and rows:
                def ts(r):
                    return datetime.datetime.strptime(r[0].strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
                inside = [r for r in rows if t_begin - 0.05 <= ts(r) <= t_end + 0.05]
                if not inside:
                    inside = sorted(rows, key=lambda r: abs(ts(r) - t_begin))[:2]
                rows = inside
            sm = sorted(float(r[1]) for r in rows)
            if sm:
                out["sm_mhz"] = sm[len(sm) // 2]
                out["sm_max_mhz"] = float(rows[0][2])
                out["samples"] = len(sm)
                reasons = set()
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                for r in rows:
                    for k, nm in enumerate(names):
                        if r[5 + k].strip().lower().startswith("active"):
                            reasons.add(nm)
                out["reasons"] = sorted(reasons)
            os.unlink(self.path)
        except Exception:
            pass
        return out


CONFIG_DESCR = {
    "c1": 'findall("\\\\d+") over synthetic ASCII text',
    "c2": 'findall("(hello|world|foo)") over a single-buffer synthetic corpus',
    "c3": "search(e-mail pattern) over synthetic 128-byte log lines (batch API)",
    "c4": 'sub("(\\\\d{3})(\\\\d{3})(\\\\d{4})", "\\\\1-\\\\2-\\\\3") over synthetic phone-number text',
    "c5": 'split("\\\\s*,\\\\s*") over a synthetic CSV corpus',
}


def default_size(config: str) -> int:
    # largest single-GPU-friendly size per config (bytes; c3: lines)
    return {"c1": 64 << 20, "c2": 1 << 30, "c3": 10_000_000, "c4": 4 << 30, "c5": 4 << 30}[config]


# --------------------------------------------------------------------------------------------
def run_reference(args):
    """CPU arm: the oracle port of the reference's path on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    import torch
    from concurrent.futures import ThreadPoolExecutor

    import oracle
    from mojo_regex_b200 import corpora

    cfg = corpora.CONFIGS[args.config]
    cores = len(os.sched_getaffinity(0))
    # bounded sample of the same corpus: 32 MiB per host thread, at most 1 GiB
    if args.config == "c3":
        sample_units = min(default_size("c3"), 250_000 * cores)
        text = corpora.make("c3", sample_units).numpy()
        sample_bytes = text.size
    else:
        sample_bytes = min(args.size or default_size(args.config), (32 << 20) * cores, GiB)
        text = corpora.make(args.config, sample_bytes).numpy()
    # one contiguous shard per thread, cut at a byte that no match can straddle
    cut_byte = 10 if args.config in ("c3", "c5") else 32
    bounds = [0]
    for k in range(1, cores):
        p = int(sample_bytes * k / cores)
        while p < sample_bytes and text[p] != cut_byte:
            p += 1
        bounds.append(min(p + 1, sample_bytes))
    bounds.append(sample_bytes)
    shards = [text[bounds[i]:bounds[i + 1]] for i in range(cores) if bounds[i + 1] > bounds[i]]
    regs = [oracle.compile(cfg["pattern"]) for _ in shards]

    def work(i):
        rx, sh = regs[i], shards[i]
        if cfg["op"] == "findall":
            return len(rx.findall(sh))
        if cfg["op"] == "sub":
            return len(rx.sub(cfg["repl"].encode(), sh))
        if cfg["op"] == "split":
            return len(rx.split_raw(sh)[1])
        return int((rx.search_batch(sh, 128)[:, 0] >= 0).sum())

    pool = ThreadPoolExecutor(max_workers=len(shards))
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        list(pool.map(work, range(len(shards))))
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    total = sum(times)
    gbs = sample_bytes * len(times) / total / 1e9
    line = {
        "impl": "reference", "metric": "input GB/s scanned", "value": gbs, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": CONFIG_DESCR[args.config], "config": args.config, "bytes_per_step": int(sample_bytes)},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": len(shards), "kind": "port",
                         "sample": "%d MiB prefix of the %s corpus, one contiguous shard per host thread; "
                                   "restated reference (oracle/), not Mojo" % (sample_bytes >> 20, args.config)},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
def cpu_baseline(config: str, cfg, shard, seconds: float = 10.0) -> dict:
    """The oracle (restated reference) on one host thread over a bounded sample of the shard."""
    import oracle

    rx = oracle.compile(cfg["pattern"])
    core = shard.buffer[shard.core_begin:shard.core_end]

    def run(sample):
        t0 = time.perf_counter()
        if cfg["op"] == "findall":
            rx.findall(sample)
        elif cfg["op"] == "sub":
            rx.sub(cfg["repl"].encode(), sample)
        elif cfg["op"] == "split":
            rx.split_raw(sample)
        else:
            rx.search_batch(sample, 128)
        return time.perf_counter() - t0

    probe_n = min(core.numel(), 8 << 20 if cfg["op"] != "search_batch" else 128 * 20000)
    probe = core[:probe_n].cpu().numpy()
    dt = max(run(probe), 1e-6)
    rate = probe_n / dt
    n = int(min(core.numel(), max(probe_n, rate * seconds / 3), 1 << 30))
    if cfg["op"] == "search_batch":
        n -= n % 128
    sample = core[:n].cpu().numpy()
    times = sorted(run(sample) for _ in range(3))
    return {"value": n / times[1] / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
            "sample": "%d MiB prefix of this rank's %s corpus, median of 3 passes, 1 thread; restated reference "
                      "(oracle/), not Mojo" % (n >> 20, config)}


# --------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local_rank: int) -> str:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned host memory is
    allocated (first touch then places the staging buffers next to the GPU's PCIe root).  Eight ranks that
    upload from unbound pinned memory measured 16 GB/s per GPU against 53 GB/s alone (round 1)."""
    try:
        import torch

        pci = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = "%04x:%02x:%02x.0" % (torch.cuda.get_device_properties(local_rank).pci_domain_id, pci,
                                    torch.cuda.get_device_properties(local_rank).pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % dom) as f:
            node = int(f.read().strip())
        if node < 0:
            return "numa node unknown"
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "numa node %d has no allowed cpus" % node
        os.sched_setaffinity(0, cpus)
        return "numa node %d (%d cpus)" % (node, len(cpus))
    except Exception as e:  # no sysfs / no permission: run unbound
        return "unbound (%s)" % type(e).__name__


L2_BYTES = 126 << 20


def measure_config(args, config: str, ctx, headline: bool) -> dict:
    """All numbers of one BASELINE config on this rank's shard; collective over the ranks (every rank calls it)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    import mojo_regex_b200 as mr
    from mojo_regex_b200 import corpora
    from mojo_regex_b200 import distributed as mrd

    world, rank, local_rank, dev, lib_stream = ctx["world"], ctx["rank"], ctx["local_rank"], ctx["dev"], ctx["lib_stream"]
    cfg = corpora.CONFIGS[config]
    size = (args.size if headline and args.size else 0) or default_size(config)
    rx = mr.compile_regex(cfg["pattern"])
    steps = args.steps if headline else max(3, min(args.steps, 10))
    warmup = args.warmup

    shard = mrd.make_shard(config, size, rank, world, dev)
    n_bytes = shard.core_bytes

    def sync_all():
        mrd.join_exchanges(lib_stream)  # queued count exchanges are issued before any other collective
        torch.cuda.synchronize()
        mr.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # inputs that fit the L2 (C1: 64 MiB) are evicted between timed steps by writing a buffer twice its size;
    # each step is then bracketed by its own pair of events and the flush stays outside
    flush = None
    if shard.buffer.numel() < 2 * L2_BYTES:
        flush = torch.empty(2 * L2_BYTES, dtype=torch.uint8, device=dev)

    # findall steps are queued without a host round trip (mrx_findall_async: the span count stays in HBM and is
    # exchanged from there); a result is released two steps later, when its kernel has long finished.  The other
    # operations return their sizes to the host by contract and are timed as they are.
    lag = collections.deque()

    def step():
        r = mrd.run_op_device(rx, cfg, shard, asynchronous=True)
        lag.append(r)
        if len(lag) > 2:
            lag.popleft().close()
        return r

    def drain():
        last = None
        while lag:
            last = lag.popleft()
            last.wait()
            if lag:
                last.close()
        return last

    clocks = ClockSampler(local_rank) if headline else None
    step()
    drain().close()  # the first result's count sizes the span arrays of the calls after it
    for _ in range(max(warmup, 3)):  # (three results in flight: the pool holds their blocks before the timed steps)
        step()
    drain().close()
    mr.set_profiling(True)
    sync_all()
    launches0 = mr.launch_count()
    t_begin = time.time()
    kernel_ms, units, alg_bytes, res = [], 0, 0, None
    if flush is None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(lib_stream):
            ev0.record(lib_stream)
            for _ in range(steps):
                step()
            mrd.join_exchanges(lib_stream)  # the count all-gathers run on their own stream: include them
            ev1.record(lib_stream)
        sync_all()
        dev_ms = ev0.elapsed_time(ev1)
        kernel_ms = mr.kernel_ms_history(min(steps, 60))
        res = drain()
        units, alg_bytes = res.units, res.algorithmic_bytes
        res.close()
    else:
        pairs = []
        for _ in range(steps):
            flush.fill_(1)
            torch.cuda.synchronize()
            with torch.cuda.stream(lib_stream):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(lib_stream)
                step()
                mrd.join_exchanges(lib_stream)
                e1.record(lib_stream)
            kernel_ms.append(mr.last_kernel_ms())
            pairs.append((e0, e1))
        sync_all()
        dev_ms = sum(a.elapsed_time(b) for a, b in pairs)
        if os.environ.get("MRX_BENCH_DEBUG"):
            print("[debug] %s per-step ms %s kernel ms %s" % (config, [round(a.elapsed_time(b), 3) for a, b in pairs],
                                                              [round(k, 3) for k in kernel_ms]), file=sys.stderr)
        res = drain()
        units, alg_bytes = res.units, res.algorithmic_bytes
        res.close()
    launches = mr.launch_count() - launches0
    t_end = time.time()
    mr.set_profiling(False)
    clk = None
    if clocks is not None:
        if t_end - t_begin < 1.0:
            # the timed region is shorter than nvidia-smi's sampling period: keep the GPU under the very same
            # kernels (untimed, and WITHOUT the count all-gather: the number of iterations is time based and
            # differs between ranks) for ~1.2 s so that the clock record is taken under load, not around it
            t_begin = time.time()
            while time.time() - t_begin < 1.2:
                mrd.run_op_local(rx, cfg, shard).close()
            torch.cuda.synchronize()
            mr.synchronize()
            t_end = time.time()
        clk = clocks.stop(t_begin, t_end)

    # ---- end to end through the public API with host buffers ---------------------------------
    host = torch.empty(shard.buffer.numel(), dtype=torch.uint8, pin_memory=True)
    host.copy_(shard.buffer)
    torch.cuda.synchronize()
    e2e_steps = max(1, min(steps, 20 if headline else 5))
    mrd.run_op_host(rx, cfg, shard, host)  # warm the pools
    sync_all()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(e2e_steps):
        d2h = mrd.run_op_host(rx, cfg, shard, host)
    mr.synchronize()
    e2e_s = time.perf_counter() - t0
    # the ceiling the host side sets: a plain upload of the same buffer from the same pinned memory, all ranks at once
    sync_all()
    scratch = torch.empty_like(shard.buffer)
    t0 = time.perf_counter()
    for _ in range(3):
        scratch.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    up_s = (time.perf_counter() - t0) / 3
    del scratch

    # ---- reduce over ranks ----------------------------------------------------------------------
    t = torch.tensor([dev_ms, e2e_s * 1e3, float(np.mean(kernel_ms)), up_s * 1e3], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(n_bytes), float(units), float(launches), float(shard.buffer.numel()), float(d2h), float(alg_bytes)],
                       dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dev_ms_max, e2e_ms_max, k_ms, up_ms = [float(x) for x in t.tolist()]
    total_bytes, total_units, total_launches, total_h2d, total_d2h, total_alg = [float(x) for x in tot.tolist()]

    # ---- CPU baseline on rank 0 (bounded sample, single thread: the reference is single-threaded)
    cpu = None
    if rank == 0 and not args.no_cpu:
        cpu = cpu_baseline(config, cfg, shard, seconds=10.0 if headline else 3.0)
    kernel = res.kernel
    flushed = flush is not None
    del host, shard, flush
    torch.cuda.empty_cache()

    peak, peak_src = ctx["peak"]
    value = total_bytes * steps / (dev_ms_max / 1e3) / 1e9
    achieved = alg_bytes / (k_ms / 1e3) / 1e9 if k_ms > 0 else None
    return {
        "value": value, "unit": "GB/s", "steps": steps, "warmup": warmup, "ms_per_step": dev_ms_max / steps,
        "config": {"workload": CONFIG_DESCR[config], "config": config,
                   "bytes_per_gpu": int(n_bytes), "units_per_step": int(total_units),
                   "l2": ("input (%.0f MiB per GPU) is larger than the 126 MB L2; no flush needed" % (n_bytes / 2**20))
                   if not flushed else "input fits the L2: a 252 MiB buffer is written between the timed steps, "
                                         "each step has its own pair of events",
                   "sharding": ("byte ranges + halo, count all-gather: " + ctx["exchange"]) if world > 1 else "single GPU"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None,
                     "traffic": (args.traffic if (args.traffic is not None and headline) else load_traffic(config, n_bytes, kernel)),
                     "kernel": kernel, "algorithmic_bytes_per_launch": int(alg_bytes), "kernel_ms": k_ms,
                     "peak_source": peak_src},
        "cpu_baseline": cpu,
        "e2e": {"value": total_bytes * e2e_steps / (e2e_ms_max / 1e3) / 1e9, "unit": "GB/s", "steps": e2e_steps,
                "h2d_bytes_per_step": int(total_h2d), "d2h_bytes_per_step": int(total_d2h),  # summed over ranks
                "upload_only_gbs": total_h2d / (up_ms / 1e3) / 1e9},
        "gpu_launches": int(total_launches),
        "clocks": clk,
    }


def run_gpu(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, "launch with torchrun --nproc-per-node N for --gpus N"
    torch.cuda.set_device(local_rank)
    binding = bind_to_gpu_numa_node(local_rank)

    import mojo_regex_b200 as mr
    from mojo_regex_b200 import distributed as mrd

    mr.init(local_rank)
    if world > 1:
        import datetime

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank),
                                timeout=datetime.timedelta(seconds=300))
    dev = torch.device("cuda", local_rank)
    exchange = "none"
    if world > 1:
        exchange = "nccl (library communicator)" if mrd.init_native_exchange(dev) else "nccl (torch.distributed)"
    ctx = {"world": world, "rank": rank, "local_rank": local_rank, "dev": dev, "exchange": exchange, "peak": load_peaks(),
           "lib_stream": torch.cuda.ExternalStream(mr.stream_handle(), device=dev)}
    wanted = ["c2", "c1", "c3", "c4", "c5"] if args.configs == "all" else [c for c in args.configs.split(",") if c]
    headline = args.config
    if headline not in wanted:
        wanted.insert(0, headline)
    results = {}
    for c in [headline] + [c for c in wanted if c != headline]:
        results[c] = measure_config(args, c, ctx, headline=(c == headline))
    if rank == 0:
        h = results[headline]
        line = {"metric": "input GB/s scanned", "value": h["value"], "unit": "GB/s", "n_gpus": world, "steps": h["steps"],
                "warmup": h["warmup"], "ms_per_step": h["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": h["config"], "roofline": h["roofline"],
                "cpu_baseline": h["cpu_baseline"], "e2e": h["e2e"], "gpu_launches": h["gpu_launches"], "clocks": h["clocks"],
                "host_binding": binding,
                "configs": {c: {k: v for k, v in r.items() if k != "clocks"} for c, r in results.items() if c != headline}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_single_process(args):
    """Evidence for the multi-device mode of the C ABI (mrx_init_multi): ONE process, ONE calling thread, N GPUs;
    every step is one host-buffer call of the public API on the whole N-shard corpus (end to end by construction:
    pinned host input, host result).  Not the driver's contract line (that is one process per GPU under torchrun);
    printed with "mode": "single-process"."""
    import numpy as np
    import torch

    import mojo_regex_b200 as mr
    from mojo_regex_b200 import corpora

    n = args.gpus
    assert "WORLD_SIZE" not in os.environ, "--single-process is not launched under torchrun"
    mr.init_multi(n)
    per_gpu = {"c1": 64 << 20, "c2": 1 << 30, "c3": 10_000_000, "c4": 1 << 30, "c5": 1 << 30}
    out = {}
    wanted = ["c2", "c1", "c3", "c4", "c5"] if args.configs == "all" else [c for c in args.configs.split(",") if c]
    for config in wanted:
        cfg = corpora.CONFIGS[config]
        parts = [corpora.make(config, per_gpu[config], seed={"c1": 0xC1, "c2": 0xC2, "c3": 0xC3, "c4": 0xC4, "c5": 0xC5}[config] + 1000003 * r,
                              device="cuda:0") for r in range(n)]
        for p_ in parts:
            if config in ("c4", "c5"):
                p_[-1] = 32 if config == "c4" else 10
        total = sum(p_.numel() for p_ in parts)
        host = torch.empty(total, dtype=torch.uint8, pin_memory=True)
        at = 0
        for p_ in parts:
            host[at:at + p_.numel()].copy_(p_)
            at += p_.numel()
        del parts
        torch.cuda.synchronize()
        rx = mr.compile_regex(cfg["pattern"])
        res = np.empty(2 * n * per_gpu["c3"], dtype=np.int64) if config == "c3" else None

        def step():
            if cfg["op"] == "findall":
                return len(rx.match_all_host(host.data_ptr(), total))
            if cfg["op"] == "search_batch":
                rx.search_batch_host(host.data_ptr(), total, res, doc_stride=128)
                return n * per_gpu["c3"]
            if cfg["op"] == "sub":
                hb = rx.sub_host(cfg["repl"], host.data_ptr(), total)
                k = hb.n
                hb.close()
                return k
            hp = rx.split_host(host.data_ptr(), total)
            k = hp.n
            hp.close()
            return k

        for _ in range(max(1, min(args.warmup, 3))):
            units = step()
        steps = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(steps):
            units = step()
        dt = (time.perf_counter() - t0) / steps
        out[config] = {"value": total / dt / 1e9, "unit": "GB/s", "ms_per_step": dt * 1e3, "bytes_per_step": int(total),
                       "units_per_step": int(units), "steps": steps}
        del host
    h = out[wanted[0]]
    print(json.dumps({"mode": "single-process", "metric": "input GB/s scanned", "value": h["value"], "unit": "GB/s", "n_gpus": n,
                      "higher_is_better": True, "scaling": "weak", "dtype": "u8", "data": "synthetic",
                      "config": {"workload": CONFIG_DESCR[wanted[0]] + " -- one mrx_*_host call per step over all GPUs "
                                             "(mrx_init_multi), host buffers in and out"},
                      "e2e": {"value": h["value"], "unit": "GB/s"}, "configs": out}))
    mr.shutdown()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c3", "c4", "c5"], help="the headline config of the line")
    ap.add_argument("--configs", default="all", help="configs measured in this run: 'all' or a comma list (the headline is "
                                                     "always included); the others go under \"configs\" of the line")
    ap.add_argument("--size", type=int, default=0, help="bytes per GPU (c3: lines); default = BASELINE size")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--single-process", action="store_true",
                    help="N GPUs from ONE process through mrx_init_multi (evidence for the C ABI's multi-device mode)")
    ap.add_argument("--traffic", type=float, default=None,
                    help="dram bytes per launch of the dominant kernel from an ncu --set full capture")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.single_process:
        run_single_process(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
