"""mojo_regex_b200/csrc/sand.hpp (Shift-And formulation of the literal-set scan, the automaton of k_sand) compiled
with g++ into a harness that walks a text exactly the way the kernel's threads do, held against the oracle wherever
the library's dense path applies."""
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
    lib.sand_scan.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_char_p, ctypes.c_int64, ctypes.c_int64,
                              ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]
    return lib


def scan(lib, lits, text: bytes, lo: int = 0, hi=None):
    """Both forms of the automaton (per-byte push, and the tail form where it fits) must agree."""
    buf = np.frombuffer(text, dtype=np.uint8) if text else np.zeros(1, np.uint8)
    cap = len(text) + 1
    hi = len(text) if hi is None else hi
    res = []
    for allow_tail in (0, 2):
        out = np.zeros((cap, 2), dtype=np.int64)
        m = lib.sand_scan(buf.ctypes.data, len(text), b"\n".join(lits), lo & ~15, lo, hi, out.ctypes.data, cap, allow_tail)
        if m == -1:
            return None
        if m != -2:
            res.append(out[:m].tolist())
    assert all(r == res[0] for r in res)
    return res[0]


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


def test_windows_tiles_and_single_literals(harness):
    """Key windows at arbitrary offsets (tile origin = lo & ~15, left context 0 / 16 / 32 bytes), texts of several
    tiles, single literals of up to 32 bytes (K = 1: the last accumulator word of a chunk is half full)."""
    rng = random.Random(5)
    text = bytes(corpora.corpus_c2(3 * 36864 + 777).numpy())
    want = oracle.compile(r"(hello|world|foo)").findall(text).tolist()
    for lo, hi in [(0, len(text)), (5, 40000), (16, 36864 + 16), (31, 36865), (33, 2 * 36864 + 48), (36864 - 2, 36864 + 3)]:
        got = scan(harness, [b"hello", b"world", b"foo"], text, lo, hi)
        assert got == [s for s in want if lo <= s[0] < hi], (lo, hi)
    import re

    for lit in [b"hello", b"a", b"zq", b"quick brown fox jump", b"the quick brown fox jumps over i"]:
        assert len(lit) <= 32
        t = bytearray(rng.choice(b"abcdefghijklmnopqrstuvwxyz ") for _ in range(40000))
        for p in (0, 130, 143, 36864 - 5, 36864 + 140, len(t) - len(lit)):
            t[p:p + len(lit)] = lit
        t = bytes(t)
        if len(lit) <= 20:      # longer operator-free literals classify MEDIUM in the reference (optimizer.mojo:481-488)
            want_l = oracle.compile(lit.decode()).findall(t).tolist()
        else:
            want_l = [list(m.span()) for m in re.finditer(re.escape(lit), t)]
        assert scan(harness, [lit], t) == want_l, lit
    assert scan(harness, [b"ab", b"cd"], b"xxabcdab") == [[2, 4], [4, 6], [6, 8]]


def test_sets_that_do_not_fit_are_refused(harness):
    assert scan(harness, [b"abcdefghijklmnopqrstuvwxyzabcdefgh"], b"x") is None     # 33 states in one slot
    assert scan(harness, [b"a%d" % i for i in range(5)], b"x") is None               # 5 literals
    assert scan(harness, [b"abcdefghi", b"jklmnopqr", b"stuvwxyza", b"bcdefghij"], b"x") is None  # 4 * 9 > 32
    assert scan(harness, [b"abcdefghi", b"jklmnopqr"], b"x") == []                   # 2 * 9 <= 32
