"""CPU check of the closed form behind the fused split kernel k_split for `X* c Y*` patterns (C5: `\\s*,\\s*`), see
DESIGN.md section 5.  The reference's required-byte findall (matcher.mojo:864-898: memchr the byte,
back-scan the first-class lookup L, one anchored DFA attempt) reduces, for the automaton S0 -X-> S1 -c-> S2(acc)
-Y-> S3(acc), to a rule per occurrence h of c that needs only membership bits and run lengths:

    start = h - (length of the maximal run of L bytes ending at h - 1)
    no span if that run holds a byte of L that is not in X (the attempt dies before it reaches c)
    end   = h + 1 + (length of the maximal run of Y bytes starting at h + 1)

Spans of neighbouring occurrences may overlap (two commas sharing one blank run): findall reports them as the
reference does, split then answers MRX_E_UNDEFINED.  Checked here against the oracle on the C5 corpus and on
random texts that contain \\v, shared runs, leading/trailing separators.  A model of the rule, not product code
(the kernel's own bit arithmetic is checked in tests/test_sepbits_cpu.py).
"""
import random

import oracle
from mojo_regex_b200 import corpora

W5 = b" \t\n\r\x0c"
W6 = W5 + b"\x0b"


def closed_form(text, c=ord(","), X=W5, Y=W5, L=W6):
    out = []
    n = len(text)
    for h in range(n):
        if text[h] != c:
            continue
        s = h
        while s > 0 and text[s - 1] in L:
            s -= 1
        if any(text[q] not in X for q in range(s, h)):
            continue
        e = h + 1
        while e < n and text[e] in Y:
            e += 1
        out.append([s, e])
    return out


def test_closed_form_equals_oracle_on_the_c5_corpus():
    rx = oracle.compile(r"\s*,\s*")
    text = bytes(corpora.make("c5", 1 << 18).numpy())
    want = rx.findall(text).tolist()
    assert closed_form(text) == want and len(want) > 10000


def test_closed_form_equals_oracle_on_random_texts():
    rx = oracle.compile(r"\s*,\s*")
    rng = random.Random(9)
    for _ in range(3000):
        t = bytes(rng.choice(b"ab,, \t\n\x0b") for _ in range(rng.randint(0, 40)))
        assert closed_form(t) == rx.findall(t).tolist(), t
