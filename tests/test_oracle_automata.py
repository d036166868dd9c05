"""Appendix-A fixtures: the exact automata the reference builds for the five configs
(SURVEY.md appendix A), plus cross-checks against Python `re` on inputs that avoid the
documented divergences (appendix B)."""
import random
import re

import pytest

import oracle
from oracle import frontend as fe

D = set(b"0123456789")
LZ = set(b"abcdefghijklmnopqrstuvwxyz")
UZ = set(b"ABCDEFGHIJKLMNOPQRSTUVWXYZ")
W5 = set(b" \t\n\r\x0c")


def edges(dfa, s):
    """-> {target: set(bytes)} for state s."""
    out = {}
    for c, t in enumerate(dfa.trans[s]):
        if t != -1:
            out.setdefault(t, set()).add(c)
    return out


def test_a1_digits_plus():
    c = fe.compile_regex(rb"\d+")
    d = c.dfa
    assert len(d.trans) == 2 and d.accepting == [False, True]
    assert edges(d, 0) == {1: D} and edges(d, 1) == {1: D}
    assert d.matcher is not None and d.simd_scan_eligible and c.required_byte == -1


def test_a2_alternation_trie():
    c = fe.compile_regex(rb"(hello|world|foo)")
    d = c.dfa
    assert len(d.trans) == 12 and d.accepting == [False, True] + [False] * 10
    assert edges(d, 1) == {}
    assert d.matcher is None and not d.is_pure_literal and c.use_pure_dfa
    for word in (b"hello", b"world", b"foo"):
        s = 0
        for ch in word:
            s = d.trans[s][ch]
            assert s != -1
        assert s == 1


def test_a3_email():
    c = fe.compile_regex(rb"[a-zA-Z0-9._%+-]+@[a-zA-Z0-9.-]+\.[a-z]{2,}")
    d = c.dfa
    C0 = LZ | UZ | D | set(b"._%+-")
    C2 = LZ | UZ | D | set(b".-")
    assert len(d.trans) == 7 and d.accepting == [False] * 6 + [True]
    assert edges(d, 0) == {1: C0}
    assert edges(d, 1) == {1: C0, 2: {ord("@")}}
    assert edges(d, 2) == {3: C2}
    assert edges(d, 3) == {3: C2 - {ord(".")}, 4: {ord(".")}}  # '.' self-loop overwritten
    assert edges(d, 4) == {5: LZ} and edges(d, 5) == {6: LZ} and edges(d, 6) == {6: LZ}
    assert c.required_byte == ord("@") and c.prefilter_literal is None
    assert d.matcher.num_ranges == 0  # 7 ranges -> nibble path
    assert not d.simd_scan_eligible


def test_a4_phone():
    c = fe.compile_regex(rb"(\d{3})(\d{3})(\d{4})")
    d = c.dfa
    assert len(d.trans) == 11 and d.accepting == [False] * 10 + [True]
    for s in range(10):
        assert edges(d, s) == {s + 1: D}
    assert edges(d, 10) == {}
    assert (c.fixed_sub_total_width, c.fixed_sub_num_groups, c.fixed_sub_concat) == (10, 3, True)
    assert c.fixed_sub_offsets[1:4] == [0, 3, 6] and c.fixed_sub_widths[1:4] == [3, 3, 4]
    assert c.required_byte == -1 and c.prefilter_literal is None


def test_a5_comma_split():
    c = fe.compile_regex(rb"\s*,\s*")
    d = c.dfa
    assert len(d.trans) == 4 and d.accepting == [False, False, True, True]
    assert edges(d, 0) == {1: W5, 2: {ord(",")}}
    assert edges(d, 1) == {1: W5, 2: {ord(",")}}
    assert edges(d, 2) == {3: W5} and edges(d, 3) == {3: W5}
    assert c.required_byte == ord(",")
    assert d.matcher.lookup[0x0B] == 1  # back-scan set carries \v, the DFA does not


def test_b1_email_multilabel_divergence():
    r = oracle.compile(r"[a-zA-Z0-9._%+-]+@[a-zA-Z0-9.-]+\.[a-z]{2,}")
    assert r.search(b"user@mail.example.com") == (0, 17)
    assert r.search(b"user@a.b.com") is None
    assert r.search(b"x user@example.com y") == (2, 18)


def test_b2_vertical_tab_skips_comma():
    r = oracle.compile(r"\s*,\s*")
    assert r.findall(b"a \x0b,b").tolist() == []
    assert r.findall(b"a ,b").tolist() == [[1, 3]]


def test_b3_overlap_refused():
    r = oracle.compile(r"\s*,\s*")
    assert r.findall(b"a,  ,b").tolist() == [[1, 4], [2, 5]]
    with pytest.raises(oracle.Unsupported):
        r.split(b"a,  ,b")


def test_required_byte_vs_dfa_path_b7():
    # sub() goes through match_next (whitespace-skip candidates), not the comma path
    r = oracle.compile(r"\s*,\s*")
    assert r.sub(b";", b"a,b ,c") == b"a,b;c"


def test_required_byte_backscan_overshoots_bounded_first_element():
    # Not in SURVEY appendix B: findall on `[A-Z]{2}-[0-9]{3}` takes the required-byte path
    # (needle '-', matcher.mojo:864-898); the back-scan runs over *all* leading [A-Z] bytes,
    # so "AZZ-110" is tried from 'A' only and fails, although search() finds (1, 7).
    r = oracle.compile(r"[A-Z]{2}-[0-9]{3}")
    assert r.compiled.required_byte == ord("-")
    assert r.findall(b"AZZ-110").tolist() == []
    assert r.search(b"AZZ-110") == (1, 7)
    assert r.findall(b"ZZ-110").tolist() == [[0, 6]]


CASES = [
    (r"\d+", "0123456789abc xyz"),
    (r"[a-z]+", "abcxyzABC019 _"),
    (r"(hello|world|foo)", "helowrdf "),
    (r"(\d{3})(\d{3})(\d{4})", "0123456789 ab"),
    (r"[a-z]+[0-9]+", "abz019 A"),
    (r"cat|dog|bird", "catdogbir "),
]


@pytest.mark.parametrize("pattern,alphabet", CASES)
def test_against_python_re_on_divergence_free_inputs(pattern, alphabet):
    rng = random.Random(hash(pattern) & 0xFFFF)
    rx = oracle.compile(pattern)
    pyre = re.compile(pattern.encode())
    for _ in range(300):
        n = rng.randrange(0, 80)
        text = bytes(ord(rng.choice(alphabet)) for _ in range(n))
        want = [[m.start(), m.end()] for m in pyre.finditer(text)]
        assert rx.findall(text).tolist() == want, (pattern, text)
        m = pyre.search(text)
        assert rx.search(text) == (m.span() if m else None)
