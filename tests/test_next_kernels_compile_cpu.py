"""The experimental kernels under mojo_regex_b200/csrc/next/ are not part of libmojoregex_cuda.so; this keeps them
compiling for sm_100a (nvcc cross-compiles without a GPU) and free of local-memory spills, so that the next round
starts from code that builds."""
import glob
import os
import re
import shutil
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCES = sorted(glob.glob(os.path.join(ROOT, "mojo_regex_b200", "csrc", "next", "*.cu")))


@pytest.mark.parametrize("src", SOURCES, ids=[os.path.basename(s) for s in SOURCES])
def test_compiles_for_sm_100a_without_spills(src):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not (os.path.exists(nvcc) or shutil.which(nvcc)):
        pytest.skip("nvcc not available")
    out = os.path.join(tempfile.mkdtemp(prefix="mrx_next_"), "k.o")
    r = subprocess.run([nvcc, "-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
                        "-Xptxas", "-v", "-c", src, "-o", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for stack, spill_st, spill_ld in re.findall(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", r.stderr):
        assert (int(stack), int(spill_st), int(spill_ld)) == (0, 0, 0), r.stderr
