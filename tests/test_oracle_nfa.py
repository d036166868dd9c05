"""The oracle's LazyDFA route (oracle/nfa.py) held against two independent computations: the direct simulation of the
PikeVM state sets (nfa.run: what PikeVMEngine._run does, pikevm.mojo:497-602) and, where leftmost-longest and
leftmost-first coincide, Python's `re`."""
import random
import re

import pytest

import oracle
from oracle import frontend as F
from oracle import nfa

PATTERNS = [r"\s+", r"(\d+)", r"\d|\w", r"(\w\d)+", r"(a|b)x", r"\w*@", r"\s", r"3[02]|40|[68]9", r"[^ab]+x",
            r"(a|ab)(c|bcd)", r"(ab)*c?", r"(?:ab) (\w+)", r"[\s\d]+x?", r"(x|y)*z{2,3}", r"a{2,}b|c"]
ALPHABET = b"ab cx1 9@yz\t_3048"


def brute_findall(prog, text):
    """LazyDFA.match_all (pikevm.mojo:779-803) on top of the direct set simulation."""
    out, pos = [], 0
    while pos <= len(text):
        m = nfa.run(prog, text, pos)
        if m:
            out.append(list(m))
            pos = max(pos + 1, m[1])
        else:
            pos += 1
    return out


@pytest.mark.parametrize("pattern", PATTERNS)
def test_subset_construction_equals_direct_simulation(pattern):
    rx = oracle.compile(pattern)
    assert rx.compiled.nfa_routed
    prog = rx.compiled.program
    rng = random.Random(hash(pattern) & 0xffff)
    for _ in range(150):
        t = bytes(rng.choice(ALPHABET) for _ in range(rng.randint(0, 40)))
        assert rx.findall(t).tolist() == brute_findall(prog, t), (pattern, t)
        for start in (0, 1, len(t) // 2, len(t)):
            want = None
            for p in range(start, len(t) + 1):   # LazyDFA.match_next (pikevm.mojo:757-777)
                want = nfa.run(prog, t, p)
                if want:
                    break
            assert rx.match_next(t, start) == want, (pattern, t, start)
            assert rx.match_first_at(t, start) == nfa.run(prog, t, start)


@pytest.mark.parametrize("pattern", [r"\s+", r"(\d+)", r"\d|\w", r"(\w\d)+", r"(a|b)x", r"\w*@", r"[^ab]+x", r"[\s\d]+x"])
def test_agrees_with_python_re_where_longest_equals_first(pattern):
    rx = oracle.compile(pattern)
    py = re.compile(pattern.replace(r"\s", r"[ \t\n\r\f]").replace(r"\d", "[0-9]").replace(r"\w", "[a-zA-Z0-9_]").encode())
    if pattern == r"[\s\d]+x":
        py = re.compile(rb"[ \t\n\r\f0-9]+x")
    if pattern == r"[^ab]+x":
        py = re.compile(rb"[^ab^]+x")   # the reference's negated class also excludes '^' (pikevm.mojo:296-333 on the raw slice)
    rng = random.Random(1)
    for _ in range(300):
        t = bytes(rng.choice(ALPHABET + b"^") for _ in range(rng.randint(0, 50)))
        assert rx.findall(t).tolist() == [list(m.span()) for m in py.finditer(t)], (pattern, t)


def test_routing_reasons():
    for pattern, word in [(r"h[ae]llo", "literal"), (r"(hello)", "literal"), (r"^[a-z]+$", "$"), (r"^(x|y)z*w", "^"),
                          (r"(a|b).*", ".*"), (r"\d*abc", "literal")]:
        with pytest.raises(oracle.Unsupported) as e:
            oracle.compile(pattern)
        assert word in str(e.value), (pattern, str(e.value))
