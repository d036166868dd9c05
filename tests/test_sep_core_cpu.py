"""mojo_regex_b200/csrc/next/sep_core.hpp (the arithmetic of the next finder for `X* c Y*` patterns; not yet
part of the library) compiled with g++ and held against the oracle: SWAR membership flags, 224-bit windows
around 144-byte chunks, run lengths by count-leading/trailing-ones."""
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
    out = os.path.join(tempfile.mkdtemp(prefix="mrx_sep_"), "libsep_harness.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out,
                           os.path.join(HERE, "native", "sep_core_harness.cpp")])
    lib = ctypes.CDLL(out)
    lib.sep_scan.restype = ctypes.c_int64
    lib.sep_scan.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                             ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_int,
                             ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_int)]
    lib.sep_detect_scan.restype = ctypes.c_int64
    lib.sep_detect_scan.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                    ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
                                    ctypes.POINTER(ctypes.c_int)]
    return lib


def scan_ws_comma(lib, text: bytes, threads: int = 256):
    """`\\s*,\\s*`: L = {9..13, 32} (simd_ops.mojo:879-883), bad = {11} (the DFA's \\s has no \\v, dfa.mojo:2901-2902)."""
    buf = np.frombuffer(text, dtype=np.uint8) if text else np.zeros(1, np.uint8)
    cap = len(text) + 1
    out = np.zeros((cap, 2), dtype=np.int64)
    extra = (ctypes.c_int * 2)(32, 0)
    bad = (ctypes.c_int * 2)(11, 0)
    ovf = ctypes.c_int(0)
    m = lib.sep_scan(buf.ctypes.data, len(text), threads, ord(","), 9, 13, extra, 1, bad, 1, out.ctypes.data, cap, ctypes.byref(ovf))
    return out[:m].tolist(), bool(ovf.value)


def longest_blank_run(text: bytes) -> int:
    best = cur = 0
    for b in text:
        cur = cur + 1 if b in b" \t\n\r\x0b\x0c" else 0
        best = max(best, cur)
    return best


def test_c5_corpus(harness):
    rx = oracle.compile(r"\s*,\s*")
    text = bytes(corpora.make("c5", 1 << 20).numpy())
    got, ovf = scan_ws_comma(harness, text)
    assert not ovf
    assert got == rx.findall(text).tolist() and len(got) > 100_000


def test_random_texts_all_chunk_alignments(harness):
    rx = oracle.compile(r"\s*,\s*")
    rng = random.Random(21)
    alphabet = b"ab,,, \t\n\r\x0b\x0c\xa0\x8c"   # incl. high bytes whose low 7 bits look like blanks
    for _ in range(1500):
        n = rng.choice([0, 1, 2, 31, 143, 144, 145, 175, 176, 177, 288, 300, 500])
        t = bytes(rng.choice(alphabet) for _ in range(n))
        for threads in (1, 2, 256):      # tiles of 144, 288 and 36 864 bytes: the halo halfwords get exercised
            got, ovf = scan_ws_comma(harness, t, threads)
            if ovf:
                assert longest_blank_run(t) >= 32
                continue
            assert got == rx.findall(t).tolist(), (threads, t)


def test_long_runs_raise_overflow_instead_of_guessing(harness):
    rx = oracle.compile(r"\s*,\s*")
    for run in (30, 31, 32, 33, 100):
        for t in (b"a" + b" " * run + b",b", b"a," + b" " * run + b"b", b" " * run + b","):
            got, ovf = scan_ws_comma(harness, t)
            if run <= 31:
                assert not ovf and got == rx.findall(t).tolist(), (run, t)
            else:
                assert ovf, (run, t)


def detect_and_scan(lib, pattern: str, text: bytes, threads: int = 2):
    """Planner (sep_plan.hpp) on the oracle's automaton for `pattern`, then the scan.  None = shape not recognised."""
    rx = oracle.compile(pattern)
    c = rx.compiled
    if c.dfa is None or c.dfa.matcher is None or c.required_byte < 0:
        return None, False, rx
    trans = np.ascontiguousarray(np.asarray(c.dfa.trans, dtype=np.int32).reshape(-1, 256))
    acc = np.asarray(c.dfa.accepting, dtype=np.uint8)
    lookup = np.asarray(c.dfa.matcher.lookup, dtype=np.uint8)
    buf = np.frombuffer(text, dtype=np.uint8) if text else np.zeros(1, np.uint8)
    cap = len(text) + 1
    out = np.zeros((cap, 2), dtype=np.int64)
    ovf = ctypes.c_int(0)
    m = lib.sep_detect_scan(trans.ctypes.data, acc.ctypes.data, trans.shape[0], c.dfa.start_state, lookup.ctypes.data,
                            c.required_byte, buf.ctypes.data, len(text), threads, out.ctypes.data, cap, ctypes.byref(ovf))
    if m == -2:
        return None, False, rx
    return out[:m].tolist(), bool(ovf.value), rx


SHAPED = [r"\s*,\s*", r"\s*;\s*", r" *, *", r"\s*,", r",\s*", r"\d*-\d*", r"[ab]*,[ab]*", r"\s*=\s*"]
NOT_SHAPED = [r"\s*,\s*,", r"[a-z]+@[a-z]+", r"\s+,\s*", r"(hello|world|foo)", r"\d+"]


def test_planner_recognises_the_family_and_the_scan_equals_the_oracle(harness):
    rng = random.Random(33)
    recognised = 0
    for pattern in SHAPED:
        try:
            got, _, rx = detect_and_scan(harness, pattern, b"")
        except (oracle.Unsupported, oracle.RegexError):
            continue
        if got is None:
            continue
        recognised += 1
        alphabet = b"ab01,;=- \t\n\x0b"
        for _ in range(300):
            t = bytes(rng.choice(alphabet) for _ in range(rng.randint(0, 320)))
            for threads in (1, 2):
                got, ovf, _ = detect_and_scan(harness, pattern, t, threads)
                if ovf:
                    continue
                assert got == rx.findall(t).tolist(), (pattern, threads, t)
    assert recognised >= 4
    for pattern in NOT_SHAPED:
        try:
            got, _, _ = detect_and_scan(harness, pattern, b"a, b")
        except (oracle.Unsupported, oracle.RegexError):
            continue
        assert got is None, pattern
