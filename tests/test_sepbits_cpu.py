"""mojo_regex_b200/csrc/sepbits.hpp (the bit-parallel separator arithmetic of the fused split kernel k_split)
compiled with g++ into a harness that splits texts chunk by chunk the way the kernel's threads do, held against the
oracle's split for patterns of the shape X* c Y*."""
import ctypes
import os
import random
import subprocess
import tempfile

import numpy as np
import pytest

import oracle
from mojo_regex_b200 import corpora

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def harness():
    out = os.path.join(tempfile.mkdtemp(prefix="mrx_sepb_"), "libsepbits_harness.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out,
                           os.path.join(HERE, "native", "sepbits_harness.cpp")])
    lib = ctypes.CDLL(out)
    lib.sepb_split.restype = ctypes.c_int64
    lib.sepb_split.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                               ctypes.POINTER(ctypes.c_int)]
    return lib


def table(c: int, L: bytes, X: bytes) -> np.ndarray:
    t = np.zeros(256, dtype=np.uint32)
    t[c] |= 1
    for b in L:
        t[b] |= 1 << 8
        if b not in X:
            t[b] |= 1 << 16
    return t


W5 = b" \t\n\r\x0c"          # the automaton's \s (dfa.mojo:2901-2902)
W6 = W5 + b"\x0b"           # the back-scan set carries \v (simd_ops.mojo:879-883)


def split(lib, text: bytes, tab: np.ndarray):
    buf = np.frombuffer(text, dtype=np.uint8) if text else np.zeros(1, np.uint8)
    out = np.zeros(len(text) + 1, dtype=np.uint8)
    offs = np.zeros(len(text) + 3, dtype=np.int64)
    flags = ctypes.c_int(0)
    p = lib.sepb_split(buf.ctypes.data, len(text), tab.ctypes.data, out.ctypes.data, offs.ctypes.data, ctypes.byref(flags))
    return out[:offs[p]].tobytes(), offs[:p + 1].tolist(), flags.value


def longest_run(text: bytes, L: bytes) -> int:
    best = cur = 0
    for b in text:
        cur = cur + 1 if b in L else 0
        best = max(best, cur)
    return best


def check(lib, pattern: str, text: bytes, tab, L):
    rx = oracle.compile(pattern)
    data, offs, flags = split(lib, text, tab)
    if flags & 1:
        assert longest_run(text, L) >= 32       # the kernel would hand this text to the general path
        return "overflow"
    try:
        wd, wo = rx.split_raw(text)
    except oracle.Unsupported:
        assert flags & 2, text
        return "overlap"
    assert not (flags & 2), text
    assert data == wd and offs == wo.tolist(), text
    return "ok"


def test_c5_corpus(harness):
    text = bytes(corpora.make("c5", 1 << 20).numpy())
    assert check(harness, r"\s*,\s*", text, table(ord(","), W6, W5), W6) == "ok"


def test_random_texts(harness):
    rng = random.Random(31)
    tab = table(ord(","), W6, W5)
    alphabet = b"ab,,, \t\n\r\x0b\x0c\xa0\x8c"
    seen = set()
    for _ in range(6000):
        n = rng.choice([0, 1, 2, 31, 47, 95, 96, 97, 143, 144, 145, 191, 192, 193, 288, 300, 500, 1000])
        # mostly texts without shared runs (those are refused): blanks only next to one comma
        if rng.random() < 0.7:
            toks = []
            while sum(map(len, toks)) < n:
                toks.append(rng.choice([b"a", b"ab", b"b" * rng.randint(1, 9), b"", b"\x0b", b"a\x0bb"]))
                toks.append(bytes(rng.choice(b" \t\n") for _ in range(rng.randint(0, 3))) + b"," +
                            bytes(rng.choice(b" \t\r\x0c") for _ in range(rng.randint(0, 3))))
            t = b"".join(toks)[:n]
        else:
            t = bytes(rng.choice(alphabet) for _ in range(n))
        seen.add(check(harness, r"\s*,\s*", t, tab, W6))
    assert seen == {"ok", "overlap"}


def test_runs_up_to_the_window_limit_and_beyond(harness):
    tab = table(ord(","), W6, W5)
    for run in list(range(0, 50)) + [63, 64, 100, 200]:
        for lead in (0, 1, 15, 16, 17, 95, 96, 143, 144, 150):
            t = b"x" * lead + b" " * run + b"," + b"\t" * run + b"yz," + b" " * (run // 2) + b"\x0b" * (run > 3) + b" ,q"
            r = check(harness, r"\s*,\s*", t, tab, W6)
            if run <= 3:
                assert r == "overlap"       # "yz, ,q": the blank belongs to both commas
            elif run <= 15:
                assert r == "ok"
            else:
                assert r in ("ok", "overflow")


def test_other_patterns_of_the_shape(harness):
    rng = random.Random(2)
    cases = [(r"\s*;\s*", ord(";"), W6, W5, b"ab; \t\n\x0b"), (r"[ab]*,[ab]*", ord(","), b"ab", b"ab", b"abc,, "),
             (r"\d*-\d*", ord("-"), b"0123456789", b"0123456789", b"019-x- "), (r"\s*=\s*", ord("="), W6, W5, b"k=v \t=")]
    for pattern, c, L, X, alphabet in cases:
        tab = table(c, L, X)
        ok = 0
        for _ in range(800):
            t = bytes(rng.choice(alphabet) for _ in range(rng.choice([0, 5, 40, 143, 144, 145, 300])))
            ok += check(harness, pattern, t, tab, L) == "ok"
        assert ok > 100
