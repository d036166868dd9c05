#!/usr/bin/env python3
"""The command profiled for the k_lines / k_lines_req captures: the C3 batch search (10 M lines x 128 B, e-mail
pattern) with the corpus resident in HBM, a few launches, nothing else.  usage: run_c3_once.py [n_lines] [reps]
(MRX_LINES_FILTER=0/1 selects the kernel); prints the average launch time from CUDA events."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mojo_regex_b200 as mr  # noqa: E402
from mojo_regex_b200 import corpora  # noqa: E402

n_lines = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda:0")
text = corpora.make("c3", n_lines, device=dev)
rx = mr.compile_regex(corpora.CONFIGS["c3"]["pattern"])
t = mr.Text.from_tensor(text)
out = torch.empty((n_lines, 2), dtype=torch.int64, device=dev)
rx.search_batch_device(t, out.data_ptr(), n_lines, doc_stride=128)
mr.set_profiling(True)
ms = []
for _ in range(reps):
    rx.search_batch_device(t, out.data_ptr(), n_lines, doc_stride=128)
    ms.append(mr.last_kernel_ms())
print("filter=%s lines=%d kernel_ms=%s matched=%.4f" % (os.environ.get("MRX_LINES_FILTER", "default"), n_lines,
      ["%.4f" % m for m in ms], (out[:, 0] >= 0).float().mean().item()))
