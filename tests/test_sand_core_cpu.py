"""mojo_regex_b200/csrc/next/sand_core.hpp (Shift-And formulation of the literal-set scan; not yet part of the
library) compiled with g++ and held against the oracle wherever the library's dense path applies."""
import ctypes
import os
import random
import subprocess
import tempfile

import numpy as np
import pytest

import mojo_regex_b200 as mr
import oracle
from mojo_regex_b200 import corpora

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def harness():
    out = os.path.join(tempfile.mkdtemp(prefix="mrx_sand_"), "libsand_harness.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out,
                           os.path.join(HERE, "native", "sand_core_harness.cpp")])
    lib = ctypes.CDLL(out)
    lib.sand_scan.restype = ctypes.c_int64
    lib.sand_scan.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int64]
    return lib


def scan(lib, lits, text: bytes):
    buf = np.frombuffer(text, dtype=np.uint8) if text else np.zeros(1, np.uint8)
    cap = len(text) + 1
    out = np.zeros((cap, 2), dtype=np.int64)
    m = lib.sand_scan(buf.ctypes.data, len(text), b"\n".join(lits), out.ctypes.data, cap)
    return None if m < 0 else out[:m].tolist()


def test_config2_literals_on_the_c2_corpus(harness):
    text = bytes(corpora.corpus_c2(1 << 20).numpy())
    got = scan(harness, [b"hello", b"world", b"foo"], text)
    want = oracle.compile(r"(hello|world|foo)").findall(text).tolist()
    assert got == want and len(got) > 10_000


def test_random_literal_sets_where_the_dense_path_applies(harness):
    rng = random.Random(13)
    done = 0
    while done < 150:
        k = rng.randint(1, 5)
        lits = sorted({bytes(rng.choice(b"abc") for _ in range(rng.randint(1, 6))) for _ in range(k)})
        pattern = lits[0].decode() if len(lits) == 1 else "(" + "|".join(x.decode() for x in lits) + ")"
        try:
            info = mr.compile_regex(pattern).info()
            orx = oracle.compile(pattern)
        except (mr.RegexError, mr.UnsupportedPattern, mr.UndefinedBehaviour, oracle.Unsupported, oracle.RegexError):
            continue
        if not (info.findall_dense and info.findall_direct):
            continue           # nesting literals: not the dense path's business; self-overlapping ones ("cc" in
                               # "ccc"): the scan yields every occurrence and the selection kernels pick, as today
        for n in (0, 1, 143, 144, 145, 600):
            t = bytes(rng.choice(b"abc ") for _ in range(n))
            got = scan(harness, lits, t)
            if got is None:
                break          # does not fit 32 state bits
            assert got == orx.findall(t).tolist(), (lits, t)
        else:
            done += 1


def test_sets_that_do_not_fit_are_refused(harness):
    assert scan(harness, [b"abcdefghijklmnopqrstuvwxyzabcdefgh"], b"x") is None     # 34 states
    assert scan(harness, [b"a%d" % i for i in range(9)], b"x") is None               # 9 literals
    assert scan(harness, [b"abcdefghi", b"jklmnopqr", b"stuvwxyza", b"bcdefghij"], b"x") is None  # 4 * 8 + 4 > 32
