#!/usr/bin/env python3
"""CPU model of the Shift-And formulation priced in DESIGN.md (round-2 candidate for k_dense): the mask
construction and the per-byte update, checked against a naive search on the C2 corpus and on random literal
sets.  Nothing here is product code.

k literals, stride s = k: literal i keeps its LAST state in bit 31 - i and the state j characters before the
end in bit 31 - i - s*j, so a shift by s moves every literal's states one character forward, the first-state
bits (INIT) can never coincide with a shifted bit (same residue class = same literal, and there is no state
before the first), i.e. (D << s) | INIT == D * 2**s + INIT (one IMAD), and "some literal ends here" is
"the top k bits of D are not all zero" -- the top BYTE of D, packed per text byte, is what k_dense's
hit extraction already consumes.  Fits when s * (max_len - 1) + k <= 32.
"""
import os
import random
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def build(lits):
    k = len(lits)
    s = k
    assert s * (max(map(len, lits)) - 1) + k <= 32, "does not fit 32 bits"
    B = [0] * 256
    init = 0
    for i, w in enumerate(lits):
        n = len(w)
        for j in range(n):  # j characters before the end
            B[w[n - 1 - j]] |= 1 << (31 - i - s * j)
        init |= 1 << (31 - i - s * (n - 1))
    return s, init, B


def scan(lits, text):
    s, init, B = build(lits)
    k = len(lits)
    D = 0
    out = []
    for pos, c in enumerate(text):
        shifted = (D << s) & 0xFFFFFFFF
        assert shifted & init == 0          # OR == ADD
        D = (shifted + init) & B[c]
        top = D >> (32 - k)
        if top:
            assert top & (top - 1) == 0     # prefix-free, nesting-free sets: one literal per end position
            i = k - 1 - (top.bit_length() - 1)
            out.append((pos + 1 - len(lits[i]), pos + 1))
    return out


def naive(lits, text):
    out = []
    for pos in range(len(text)):
        for w in lits:
            if text[pos:pos + len(w)] == w:
                out.append((pos, pos + len(w)))
    return sorted(out, key=lambda t: t[1])


if __name__ == "__main__":
    from mojo_regex_b200 import corpora
    lits = [b"hello", b"world", b"foo"]
    text = bytes(corpora.corpus_c2(1 << 18).numpy())
    got = scan(lits, text)
    assert got == naive(lits, text) and len(got) > 1000
    print("C2 corpus: %d occurrences, Shift-And == naive" % len(got))
    rng = random.Random(5)
    done = 0
    while done < 200:
        k = rng.randint(1, 5)
        ls = list({bytes(rng.choice(b"abc") for _ in range(rng.randint(1, 6))) for _ in range(k)})
        # the dense path's precondition (program.cpp: build_dense): no literal inside another one
        if any(a != b and a in b for a in ls for b in ls):
            continue
        if len(ls) * (max(map(len, ls)) - 1) + len(ls) > 32:
            continue
        t = bytes(rng.choice(b"abc ") for _ in range(400))
        assert scan(ls, t) == naive(ls, t), (ls, t)
        done += 1
    print("200 random literal sets: Shift-And == naive")
