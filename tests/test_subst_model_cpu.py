"""CPU model of the arithmetic of k_subst (mojo_regex_b200/csrc/k_subst.cu, fused sub for fixed-width chains over one
class, BASELINE config 4), chunk by chunk as the kernel's threads do it, held against the oracle's sub:

    E      = positions followed by W class bytes            (AND-smear by doubling over the chunk's 144 bits + 32 after)
    seeds  = run starts in E; for a run that continues from before the chunk, the first start at the run's phase
             (carry % W), when W more bytes of the run remain
    M      = closure of seeds under  M |= (M << W) & E      (a hop stays inside its run: both ends see W class bytes)
    skip   = bytes at the chunk start that belong to a match begun in the previous chunk
    output = bytes outside matches at position + (R - W) * (starts before), the template at every start

The kernel itself is checked on the GPU (tests/test_gpu_parity.py: fused vs general path vs oracle); this file pins
the closed form it implements.  Bit masks are Python integers, bit i = byte i of the chunk."""
import random

import numpy as np
import pytest

import oracle

CHUNK, TILE = 144, 36864


def and_smear(x, w):
    p = 1
    while 2 * p <= w:
        x &= x >> p
        p *= 2
    if w > p:
        x &= x >> (w - p)
    return x


def or_smear(x, w):
    p = 1
    while 2 * p <= w:
        x |= x << p
        p *= 2
    if w > p:
        x |= x << (w - p)
    return x


def chunk_matches(bits, nxt32, carry, w):
    """-> (mask of match starts in the chunk, skip) from the chunk's class bits, the 32 class bits after it and the
    number of class bytes immediately before it -- nothing else, exactly what a kernel thread has."""
    own = (1 << CHUNK) - 1
    window = bits | (nxt32 << CHUNK)
    e = and_smear(window, w) & own
    before = (bits << 1) | (1 if carry > 0 else 0)
    m = bits & ~before & e
    skip = 0
    if carry > 0 and bits & 1:
        lead = 0
        while lead < CHUNK and (bits >> lead) & 1:
            lead += 1
        if lead == CHUNK:
            ext = 0
            while ext < 32 and (nxt32 >> ext) & 1:
                ext += 1
            lead += ext
        ph = carry % w
        r0 = w - ph if ph else 0
        if ph and lead >= r0:
            skip = r0
        if r0 + w <= lead:
            m |= 1 << r0
    while True:
        t = (m << w) & e & ~m & own
        if not t:
            break
        m |= t
    return m, skip


def model_sub(text: bytes, cls: bytes, w: int, template):
    """template: list of ints (offset into the match) / bytes objects (literals)."""
    member = bytes(1 if b in cls else 0 for b in range(256))
    n = len(text)
    out = bytearray()
    carry = 0
    for c0 in range(0, n, CHUNK):
        chunk = text[c0:c0 + CHUNK]
        bits = 0
        for i, b in enumerate(chunk):
            bits |= member[b] << i
        nxt = 0
        for i, b in enumerate(text[c0 + CHUNK:c0 + CHUNK + 32]):
            nxt |= member[b] << i
        m, skip = chunk_matches(bits, nxt, carry, w)
        cov = or_smear(m, w) | ((1 << skip) - 1)
        for i, b in enumerate(chunk):
            if (m >> i) & 1:
                for t in template:
                    out += bytes([text[c0 + i + t]]) if isinstance(t, int) else t
            elif not (cov >> i) & 1:
                out.append(b)
        # run length entering the next chunk
        trail = 0
        while trail < len(chunk) and member[chunk[len(chunk) - 1 - trail]]:
            trail += 1
        carry = carry + trail if trail == len(chunk) else trail
    return bytes(out)


CASES = [
    (r"(\d{3})(\d{3})(\d{4})", rb"\1-\2-\3", b"0123456789", 10, [0, 1, 2, b"-", 3, 4, 5, b"-", 6, 7, 8, 9]),
    (r"(\d{3})(\d{3})(\d{4})", rb"\3\1", b"0123456789", 10, [6, 7, 8, 9, 0, 1, 2]),
    (r"\d{4}", rb"<#>", b"0123456789", 4, [b"<#>"]),
    (r"[a-c]{32}", rb"_", b"abc", 32, [b"_"]),
    (r"[ab]{5}", rb"", b"ab", 5, []),
]


@pytest.mark.parametrize("pattern,repl,cls,w,template", CASES)
def test_model_equals_the_oracle(pattern, repl, cls, w, template):
    rx = oracle.compile(pattern)
    rng = random.Random(w * 7 + len(repl))
    other = b" xyz-,."
    texts = [b"", cls[:1] * (w - 1), cls[:1] * w, cls[:1] * (3 * CHUNK + 5), b"." * 143 + cls[:1] * (2 * w + 3)]
    parts = []
    for _ in range(6000):          # runs of every length around the width and its multiples
        parts.append(bytes(rng.choice(cls) for _ in range(rng.randrange(0, 4 * w + 3))))
        parts.append(bytes(rng.choice(other) for _ in range(rng.randrange(1, 4))))
    texts.append(b"".join(parts))
    edge = bytearray(b"." * (2 * TILE + 300))   # runs across chunk and tile edges, at every phase
    for at in (CHUNK, 7 * CHUNK, TILE, 2 * TILE):
        for back in (1, w - 1, w, 2 * w + 3):
            edge[at - back:at - back + 3 * w + 1] = bytes(rng.choice(cls) for _ in range(3 * w + 1))
            texts.append(bytes(edge))
            edge[at - back:at - back + 3 * w + 1] = b"." * (3 * w + 1)
    giant = bytearray(cls[:1] * (TILE + 1000))  # one run over a tile edge, started off phase
    giant[13] = ord("x")
    texts.append(bytes(giant))
    for text in texts:
        want = rx.sub(repl, np.frombuffer(text, np.uint8)) if text else b""
        assert model_sub(text, cls, w, template) == want, (pattern, len(text))
