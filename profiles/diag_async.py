"""Diagnostic (not part of the product): host and device time of asynchronous findall steps, as bench.py issues them."""
import collections, sys, time
import torch
import mojo_regex_b200 as mr
from mojo_regex_b200 import corpora, distributed as mrd

config = sys.argv[1] if len(sys.argv) > 1 else "c2"
prof = len(sys.argv) > 2 and sys.argv[2] == "prof"
mr.init(0)
dev = torch.device("cuda:0")
cfg = corpora.CONFIGS[config]
shard = mrd.make_shard(config, cfg["full_bytes"], 0, 1, dev)
rx = mr.compile_regex(cfg["pattern"])
lib_stream = torch.cuda.ExternalStream(mr.stream_handle()) if hasattr(mr, "stream_handle") else None
lag = collections.deque()
def step():
    t0 = time.perf_counter()
    r = mrd.run_op_device(rx, cfg, shard, asynchronous=True)
    t1 = time.perf_counter()
    lag.append(r)
    if len(lag) > 2:
        lag.popleft().close()
    return (t1 - t0) * 1e3, (time.perf_counter() - t1) * 1e3
for _ in range(4): step()
mr.synchronize()
mr.set_profiling(prof)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for mode in ("back-to-back", "flushed"):
    T0 = time.perf_counter(); times = []
    for _ in range(10):
        if mode == "flushed":
            flush.fill_(1); torch.cuda.synchronize()
        times.append(step())
    mr.synchronize(); T1 = time.perf_counter()
    print(config, mode, "prof" if prof else "", "wall per step ms %.3f" % ((T1 - T0) / 10 * 1e3), "call ms", [round(a, 3) for a, b in times][:6], "close ms", [round(b, 3) for a, b in times][:6])
    if prof: print("kernel ms", [round(x, 4) for x in mr.kernel_ms_history(10)])
# as bench.py's flushed loop: one event pair per step on the library stream
lib_stream = torch.cuda.ExternalStream(mr.stream_handle(), device=dev)
pairs = []
for _ in range(8):
    flush.fill_(1); torch.cuda.synchronize()
    with torch.cuda.stream(lib_stream):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(lib_stream)
        step()
        mrd.join_exchanges(lib_stream)
        e1.record(lib_stream)
    pairs.append((e0, e1))
torch.cuda.synchronize(); mr.synchronize()
print("event pairs ms", [round(a.elapsed_time(b), 3) for a, b in pairs])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(lib_stream):
    e0.record(lib_stream)
    for _ in range(10): step()
    mrd.join_exchanges(lib_stream)
    e1.record(lib_stream)
torch.cuda.synchronize(); mr.synchronize()
print("back-to-back by events ms/step", round(e0.elapsed_time(e1) / 10, 4))
