"""oracle/scan.c: the AVX2 forms of find_first_nibble_match / count_consecutive_matches (what bench.py's CPU
baseline times) against the scalar loops (ORACLE_NO_AVX2=1), on every op that goes through them."""
import numpy as np
import pytest

import oracle

PATTERNS = [r"\d+", r"[a-z]+", r"(hello|world|foo)", r"(\d{3})(\d{3})(\d{4})", r"[a-zA-Z0-9._%+-]+@[a-zA-Z0-9.-]+\.[a-zA-Z]{2,}",
            r"[\x80-\xff]+", r"[0-9a-fx]+", r"[a-c]{4}", r"\w+", r"\s+", r"[^a]+", r"[a-z]+[0-9]+"]


def _both(pattern, monkeypatch, simd_width=32):
    monkeypatch.delenv("ORACLE_NO_AVX2", raising=False)
    fast = oracle.compile(pattern, simd_width)
    monkeypatch.setenv("ORACLE_NO_AVX2", "1")
    slow = oracle.compile(pattern, simd_width)
    monkeypatch.delenv("ORACLE_NO_AVX2", raising=False)
    return fast, slow


@pytest.mark.parametrize("pattern", PATTERNS)
@pytest.mark.parametrize("simd_width", [16, 32])
def test_avx2_scans_equal_the_scalar_loops(pattern, simd_width, monkeypatch):
    fast, slow = _both(pattern, monkeypatch, simd_width)
    rng = np.random.default_rng(len(pattern) * 31 + simd_width)
    alphabets = [np.arange(256, dtype=np.uint8), np.frombuffer(b"ab019 .@xyzhelowrdf\xa0\xff", dtype=np.uint8),
                 np.frombuffer(b"0123456789", dtype=np.uint8), np.frombuffer(b"a ", dtype=np.uint8)]
    for alphabet in alphabets:
        for n in (0, 1, 15, 16, 17, 31, 32, 33, 63, 64, 65, 95, 1000, 40_000):
            text = alphabet[rng.integers(0, len(alphabet), n)].copy()
            assert np.array_equal(fast.findall(text), slow.findall(text)), (pattern, n)
            for start in (0, 1, 7, 33):
                assert fast.match_next(text, start) == slow.match_next(text, start)
                assert fast.match_first_at(text, start) == slow.match_first_at(text, start)
    text = alphabets[1][rng.integers(0, len(alphabets[1]), 5000)].copy()
    assert fast.sub(b"<>", text) == slow.sub(b"<>", text)


def test_long_runs_cross_many_vectors(monkeypatch):
    fast, slow = _both(r"\d+", monkeypatch)
    for run in (31, 32, 33, 64, 100, 1000):
        for lead in (0, 1, 31, 32, 33):
            text = np.frombuffer(b"x" * lead + b"7" * run + b"yz" + b"1" * run, dtype=np.uint8)
            assert fast.findall(text).tolist() == slow.findall(text).tolist() == [[lead, lead + run], [lead + run + 2, lead + 2 * run + 2]]
