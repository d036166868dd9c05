"""Seeded synthetic corpora for the five BASELINE.json configs.

Vectorised torch code so the same generator runs on the CPU (tests, small sizes) and on
the GPU (bench, full sizes).  Every corpus is a flat uint8 tensor.  Generators are
deterministic for a given (size, seed, device type).
"""

from __future__ import annotations

from typing import Tuple

import torch

LOWER = "abcdefghijklmnopqrstuvwxyz"
ALNUM_ = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789_"
LOCAL = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789._%+-"
LABEL = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789-"


def _gen(seed: int, device) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return g


def _table(chars: str, device) -> torch.Tensor:
    return torch.tensor(list(chars.encode()), dtype=torch.uint8, device=device)


def _ragged(lengths: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """-> (token id per byte, offset inside the token per byte, total bytes)."""
    starts = torch.cumsum(lengths, 0) - lengths
    total = int(lengths.sum().item())
    tok = torch.repeat_interleave(torch.arange(lengths.numel(), device=lengths.device), lengths, output_size=total)
    off = torch.arange(total, device=lengths.device) - starts[tok]
    return tok, off, total


def _fit(buf: torch.Tensor, n: int, pad: int = 32) -> torch.Tensor:
    if buf.numel() >= n:
        return buf[:n].contiguous()
    out = torch.full((n,), pad, dtype=torch.uint8, device=buf.device)
    out[: buf.numel()] = buf
    return out


def _chunked(fn, n: int, seed: int, device, chunk: int = 1 << 28) -> torch.Tensor:
    """Generate big corpora in independent 256 MiB pieces (bounded temporary memory)."""
    if n <= chunk:
        return fn(n, seed, device)
    parts, done, k = [], 0, 0
    while done < n:
        m = min(chunk, n - done)
        parts.append(fn(m, seed + 7919 * k, device))
        done += m
        k += 1
    return torch.cat(parts)


# ---- C1: findall(r"\d+") -- words / digit runs / single spaces, ~27 % digits, ~1 match / 17 B -------
def _c1(n: int, seed: int, device) -> torch.Tensor:
    g = _gen(seed, device)
    n_tok = n // 5 + 16
    is_digit = torch.rand(n_tok, generator=g, device=device) < 0.40
    wlen = torch.randint(1, 13, (n_tok,), generator=g, device=device)
    dlen = torch.randint(1, 9, (n_tok,), generator=g, device=device)
    lengths = torch.where(is_digit, dlen, wlen) + 1  # + trailing space
    tok, off, total = _ragged(lengths)
    r = torch.randint(0, 260, (total,), generator=g, device=device)
    letters = (r % 26 + 97).to(torch.uint8)
    digits = (r % 10 + 48).to(torch.uint8)
    out = torch.where(is_digit[tok], digits, letters)
    out[off == lengths[tok] - 1] = 32
    return _fit(out, n)


def corpus_c1(n: int = 64 << 20, seed: int = 0xC1, device="cpu") -> torch.Tensor:
    return _chunked(_c1, n, seed, device)


# ---- C2: findall(r"(hello|world|foo)") -- dictionary words, hits ~1 / 64 B, near-misses ------------
def _dictionary(device) -> Tuple[torch.Tensor, torch.Tensor]:
    g = torch.Generator(device="cpu")
    g.manual_seed(0xD1C7)
    words = []
    while len(words) < 4096:
        ln = int(torch.randint(2, 11, (1,), generator=g))
        w = bytes((torch.randint(0, 26, (ln,), generator=g) + 97).tolist())
        if b"hello" in w or b"world" in w or b"foo" in w:
            continue
        words.append(w)
    special = [b"hello", b"world", b"foo", b"hell", b"worl", b"fo", b"helloworld", b"foofoo"]
    allw = words + special
    mat = torch.full((len(allw), 12), 32, dtype=torch.uint8)
    lens = torch.zeros(len(allw), dtype=torch.int64)
    for i, w in enumerate(allw):
        mat[i, : len(w)] = torch.tensor(list(w), dtype=torch.uint8)
        lens[i] = len(w)
    return mat.to(device), lens.to(device)


def _c2(n: int, seed: int, device) -> torch.Tensor:
    g = _gen(seed, device)
    mat, lens = _dictionary(device)
    n_tok = n // 5 + 16
    u = torch.rand(n_tok, generator=g, device=device)
    wid = torch.randint(0, 4096, (n_tok,), generator=g, device=device)
    hit = torch.randint(4096, 4099, (n_tok,), generator=g, device=device)
    miss = torch.randint(4099, 4104, (n_tok,), generator=g, device=device)
    wid = torch.where(u < 0.105, hit, wid)
    wid = torch.where((u >= 0.105) & (u < 0.125), miss, wid)
    lengths = lens[wid] + 1
    tok, off, total = _ragged(lengths)
    out = mat[wid[tok], off.clamp(max=11)]
    out[off == lengths[tok] - 1] = 32
    return _fit(out, n)


def corpus_c2(n: int = 1 << 30, seed: int = 0xC2, device="cpu") -> torch.Tensor:
    return _chunked(_c2, n, seed, device)


# ---- C3: search(e-mail) over 128-byte log lines ---------------------------------------------------
def _c3(n_lines: int, seed: int, device) -> torch.Tensor:
    g = _gen(seed, device)
    W = 128
    r = torch.randint(0, 26 * 7, (n_lines, W), generator=g, device=device)
    lines = (r % 26 + 97).to(torch.uint8)
    lines[(r // 26) == 0] = 32  # ~1/7 spaces -> words of ~6 letters
    u = torch.rand(n_lines, generator=g, device=device)
    has_addr = u < 0.30
    decoy = (u >= 0.30) & (u < 0.35)
    # address geometry
    a = torch.randint(1, 21, (n_lines,), generator=g, device=device)       # local part
    nl = torch.randint(1, 4, (n_lines,), generator=g, device=device)      # domain labels
    l = torch.randint(1, 9, (n_lines, 3), generator=g, device=device)     # label lengths
    tl = torch.randint(2, 7, (n_lines,), generator=g, device=device)      # TLD
    l = l * (torch.arange(3, device=device)[None, :] < nl[:, None])
    dom = l.sum(1) + (nl - 1)                                              # labels joined by '.'
    alen = a + 1 + dom + 1 + tl
    off = (torch.rand(n_lines, generator=g, device=device) * (W - 3 - alen).clamp(min=1)).long() + 1
    col = torch.arange(W, device=device)[None, :] - off[:, None]           # position inside the address
    rc = torch.randint(0, 1 << 20, (n_lines, W), generator=g, device=device)
    loc_t, lab_t = _table(LOCAL, device), _table(LABEL, device)
    addr = torch.zeros((n_lines, W), dtype=torch.uint8, device=device)
    in_local = (col >= 0) & (col < a[:, None])
    addr = torch.where(in_local, loc_t[rc % len(LOCAL)], addr)
    at_pos = col == a[:, None]
    addr = torch.where(at_pos, torch.tensor(64, dtype=torch.uint8, device=device), addr)
    dcol = col - a[:, None] - 1                                            # position inside the domain
    in_dom = (dcol >= 0) & (dcol < dom[:, None])
    b1 = l[:, 0:1]
    b2 = b1 + 1 + l[:, 1:2]
    dot = ((dcol == b1) & (nl[:, None] >= 2)) | ((dcol == b2) & (nl[:, None] >= 3))
    addr = torch.where(in_dom, lab_t[rc % len(LABEL)], addr)
    addr = torch.where(in_dom & dot, torch.tensor(46, dtype=torch.uint8, device=device), addr)
    tcol = dcol - dom[:, None]
    addr = torch.where(tcol == 0, torch.tensor(46, dtype=torch.uint8, device=device), addr)
    in_tld = (tcol >= 1) & (tcol <= tl[:, None])
    addr = torch.where(in_tld, (rc % 26 + 97).to(torch.uint8), addr)
    inside = (col >= 0) & (col < alen[:, None])
    border = (col == -1) | (col == alen[:, None])
    sel = has_addr[:, None]
    lines = torch.where(sel & inside, addr, lines)
    lines = torch.where(sel & border, torch.tensor(32, dtype=torch.uint8, device=device), lines)
    # decoys: "name@host " without a dot, or a one-letter TLD "n@h.c "
    dsel = decoy[:, None]
    dlen = a.clamp(max=8)
    one_tld = (rc[:, 0] & 1) == 1
    dcolx = col
    d_addr = torch.where(dcolx < dlen[:, None], (rc % 26 + 97).to(torch.uint8), torch.zeros_like(addr))
    d_addr = torch.where(dcolx == dlen[:, None], torch.tensor(64, dtype=torch.uint8, device=device), d_addr)
    d_addr = torch.where((dcolx > dlen[:, None]) & (dcolx <= dlen[:, None] + 4), (rc % 26 + 97).to(torch.uint8), d_addr)
    d_addr = torch.where((dcolx == dlen[:, None] + 5) & one_tld[:, None], torch.tensor(46, dtype=torch.uint8, device=device), d_addr)
    d_addr = torch.where((dcolx == dlen[:, None] + 6) & one_tld[:, None], (rc % 26 + 97).to(torch.uint8), d_addr)
    d_total = dlen + 5 + torch.where(one_tld, 2, 0)
    d_inside = (dcolx >= 0) & (dcolx < d_total[:, None])
    d_border = (dcolx == -1) | (dcolx == d_total[:, None])
    lines = torch.where(dsel & d_inside, d_addr, lines)
    lines = torch.where(dsel & d_border, torch.tensor(32, dtype=torch.uint8, device=device), lines)
    lines[:, W - 1] = 10
    return lines.reshape(-1)


def corpus_c3(n_lines: int = 10_000_000, seed: int = 0xC3, device="cpu") -> torch.Tensor:
    chunk = 1 << 20
    if n_lines <= chunk:
        return _c3(n_lines, seed, device)
    parts, done, k = [], 0, 0
    while done < n_lines:
        m = min(chunk, n_lines - done)
        parts.append(_c3(m, seed + 7919 * k, device))
        done += m
        k += 1
    return torch.cat(parts)


# ---- C4: sub(phone) -- filler words + digit runs of length 10 (60 %), 7/9/11/12 (30 %), 20/25/30 (10 %)
def _c4(n: int, seed: int, device) -> torch.Tensor:
    g = _gen(seed, device)
    n_tok = n // 6 + 16
    is_run = torch.rand(n_tok, generator=g, device=device) < 0.217
    wlen = torch.randint(1, 13, (n_tok,), generator=g, device=device)
    u = torch.rand(n_tok, generator=g, device=device)
    pick = torch.randint(0, 12, (n_tok,), generator=g, device=device)
    mid = torch.tensor([7, 9, 11, 12], device=device)[pick % 4]
    big = torch.tensor([20, 25, 30], device=device)[pick % 3]
    rlen = torch.where(u < 0.6, torch.full_like(mid, 10), torch.where(u < 0.9, mid, big))
    lengths = torch.where(is_run, rlen, wlen) + 1
    tok, off, total = _ragged(lengths)
    r = torch.randint(0, 260, (total,), generator=g, device=device)
    out = torch.where(is_run[tok], (r % 10 + 48).to(torch.uint8), (r % 26 + 97).to(torch.uint8))
    out[off == lengths[tok] - 1] = 32
    return _fit(out, n)


def corpus_c4(n: int = 4 << 30, seed: int = 0xC4, device="cpu") -> torch.Tensor:
    return _chunked(_c4, n, seed, device)


# ---- C5: split(r"\s*,\s*") -- CSV rows of 8-16 fields -------------------------------------------------
def _c5(n: int, seed: int, device) -> torch.Tensor:
    g = _gen(seed, device)
    n_rows = n // 60 + 4
    per_row = torch.randint(8, 17, (n_rows,), generator=g, device=device)
    row_of, idx_in_row, n_fields = _ragged(per_row)
    first = idx_in_row == 0
    last = idx_in_row == per_row[row_of] - 1
    flen = torch.randint(0, 13, (n_fields,), generator=g, device=device)
    flen = torch.where((first | last) & (flen == 0), torch.ones_like(flen), flen)
    empty = flen == 0
    ws_b = torch.randint(0, 3, (n_fields,), generator=g, device=device) * (torch.rand(n_fields, generator=g, device=device) < 0.5)
    ws_a = torch.randint(0, 3, (n_fields,), generator=g, device=device) * (torch.rand(n_fields, generator=g, device=device) < 0.5)
    # separator i sits after field i (none after the last field of a row, which gets '\n')
    nxt_empty = torch.cat([empty[1:], torch.zeros(1, dtype=torch.bool, device=device)])
    ws_b = torch.where(empty | last, torch.zeros_like(ws_b), ws_b)
    ws_a = torch.where(nxt_empty | last, torch.zeros_like(ws_a), ws_a)
    lengths = flen + ws_b + 1 + ws_a  # field, ws, ',' (or '\n'), ws
    tok, off, total = _ragged(lengths)
    r = torch.randint(0, 63 * 4, (total,), generator=g, device=device)
    tbl = _table(ALNUM_, device)
    out = tbl[r % 63]
    f, b = flen[tok], ws_b[tok]
    is_ws = ((off >= f) & (off < f + b)) | (off > f + b)
    out = torch.where(is_ws, torch.where(r % 2 == 0, torch.tensor(32, dtype=torch.uint8, device=device),
                                         torch.tensor(9, dtype=torch.uint8, device=device)), out)
    sep = off == f + b
    out = torch.where(sep, torch.where(last[tok], torch.tensor(10, dtype=torch.uint8, device=device),
                                       torch.tensor(44, dtype=torch.uint8, device=device)), out)
    return _fit(out, n, pad=120)


def corpus_c5(n: int = 32 << 30, seed: int = 0xC5, device="cpu") -> torch.Tensor:
    return _chunked(_c5, n, seed, device)


CONFIGS = {
    "c1": dict(op="findall", pattern=r"\d+", full_bytes=64 << 20),
    "c2": dict(op="findall", pattern=r"(hello|world|foo)", full_bytes=1 << 30),
    "c3": dict(op="search_batch", pattern=r"[a-zA-Z0-9._%+-]+@[a-zA-Z0-9.-]+\.[a-z]{2,}", full_lines=10_000_000),
    "c4": dict(op="sub", pattern=r"(\d{3})(\d{3})(\d{4})", repl=r"\1-\2-\3", full_bytes=4 << 30),
    "c5": dict(op="split", pattern=r"\s*,\s*", full_bytes=32 << 30),
}


def make(config: str, size: int, seed=None, device="cpu") -> torch.Tensor:
    """size = bytes (c1, c2, c4, c5) or lines (c3)."""
    fn = {"c1": corpus_c1, "c2": corpus_c2, "c3": corpus_c3, "c4": corpus_c4, "c5": corpus_c5}[config]
    return fn(size, device=device) if seed is None else fn(size, seed=seed, device=device)
