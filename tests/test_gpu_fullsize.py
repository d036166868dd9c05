"""Parity at BASELINE.json's full single-GPU sizes, through an INDEPENDENT computation of the same answer.

The oracle (C restatement) is too slow for gigabytes; here every expected result is derived from the corpus
with plain torch tensor operations on the GPU (boolean masks, shifts, cumulative sums) -- no regex code
involved -- and compared bit for bit with what the CUDA path returns through the C ABI:

  C1  findall(\\d+), 64 MiB      spans == (run starts, run ends) of the digit mask
  C2  findall((hello|world|foo)), 1 GiB   spans == every occurrence of the three literals (they cannot overlap)
  C3  search(e-mail) x 10 M lines  k_lines == k_scan (two unrelated GPU algorithms) + the oracle on a sample
  C4  sub(phone, \\1-\\2-\\3), 4 GiB   removing every '-' gives the input back; #'-' == 2 * sum(floor(run/10))
  C5  split(\\s*,\\s*), 4 GiB     pieces == the text without commas and the whitespace runs touching them;
      offsets == kept bytes before each comma
"""
import os

import numpy as np
import pytest
import torch

import mojo_regex_b200 as mr
import oracle
from mojo_regex_b200 import corpora

pytestmark = pytest.mark.gpu
CHUNK = 1 << 28


def _dev():
    return torch.device("cuda", 0)


def _shifted_prev(mask: torch.Tensor, first: bool) -> torch.Tensor:
    """mask shifted right by one position; `first` is the value before element 0."""
    out = torch.empty_like(mask)
    out[1:] = mask[:-1]
    out[0] = first
    return out


def test_c1_full_size_runs_equal_mask_transitions():
    n = 64 << 20
    text = corpora.make("c1", n, device=_dev())
    d = (text >= 48) & (text <= 57)
    prev = _shifted_prev(d, False)
    nxt = torch.empty_like(d)
    nxt[:-1] = d[1:]
    nxt[-1] = False
    starts = torch.nonzero(d & ~prev).flatten()
    ends = torch.nonzero(d & ~nxt).flatten() + 1
    ds = mr.compile_regex(r"\d+").match_all_device(mr.Text.from_tensor(text))
    got = ds.tensor()
    assert got.shape[0] == starts.numel() == 4_009_809 or got.shape[0] == starts.numel()
    assert torch.equal(got[:, 0], starts) and torch.equal(got[:, 1], ends)
    ds.close()


def test_c2_full_size_spans_equal_all_literal_occurrences():
    n = 1 << 30
    text = corpora.make("c2", n, device=_dev())
    t = mr.Text.from_tensor(text)
    ds = mr.compile_regex(r"(hello|world|foo)").match_all_device(t)
    got = ds.tensor()
    exp_s, exp_e = [], []
    pad = torch.zeros(8, dtype=torch.uint8, device=text.device)
    for c0 in range(0, n, CHUNK):
        c1 = min(n, c0 + CHUNK)
        w = torch.cat([text[c0:min(n, c1 + 8)], pad])  # window with look-ahead for literals that straddle the chunk
        m = c1 - c0
        hit_len = torch.zeros(m, dtype=torch.int64, device=text.device)
        for lit in (b"hello", b"world", b"foo"):
            ok = torch.ones(m, dtype=torch.bool, device=text.device)
            for k, ch in enumerate(lit):
                ok &= w[k:k + m] == ch
            hit_len[ok] = len(lit)
        pos = torch.nonzero(hit_len).flatten()
        exp_s.append(pos + c0)
        exp_e.append(pos + c0 + hit_len[pos])
        del hit_len, ok
    exp_s, exp_e = torch.cat(exp_s), torch.cat(exp_e)
    keep = exp_e <= n
    exp_s, exp_e = exp_s[keep], exp_e[keep]
    assert got.shape[0] == exp_s.numel()
    assert torch.equal(got[:, 0], exp_s) and torch.equal(got[:, 1], exp_e)
    ds.close()


def test_c3_full_size_two_gpu_algorithms_agree_and_sample_matches_oracle():
    n_lines = 10_000_000
    text = corpora.make("c3", n_lines, device=_dev())
    pat = corpora.CONFIGS["c3"]["pattern"]
    rx = mr.compile_regex(pat)
    t = mr.Text.from_tensor(text)
    a = torch.empty((n_lines, 2), dtype=torch.int64, device=text.device)
    b = torch.empty((n_lines, 2), dtype=torch.int64, device=text.device)
    assert rx.info().search_rev_states > 0
    rx.search_batch_device(t, a.data_ptr(), n_lines, doc_stride=128)          # k_lines (reverse automaton)
    os.environ["MRX_NO_LINES"] = "1"
    try:
        rx.search_batch_device(t, b.data_ptr(), n_lines, doc_stride=128)      # k_scan (candidate queue)
    finally:
        del os.environ["MRX_NO_LINES"]
    assert torch.equal(a, b)
    frac = (a[:, 0] >= 0).float().mean().item()
    assert 0.05 < frac < 0.5   # 30 % of the lines carry an address; multi-label domains fail in the reference (SURVEY B-1)
    orx = oracle.compile(pat)
    rng = np.random.default_rng(11)
    idx = rng.integers(0, n_lines, 20_000)
    host = a[torch.from_numpy(idx).to(text.device)].cpu().numpy()
    lines = text.view(n_lines, 128)[torch.from_numpy(idx).to(text.device)].cpu().numpy()
    for k in range(idx.size):
        r = orx.search(lines[k])
        assert tuple(host[k]) == (r if r else (-1, -1))


def test_c4_full_size_sub_round_trip_and_dash_count():
    n = 4 << 30
    text = corpora.make("c4", n, device=_dev())
    assert int((text == 45).sum().item()) == 0          # the corpus holds no '-'
    # expected number of matches: floor(run length / 10) summed over the digit runs
    n_matches = 0
    carry = 0      # length of the digit run that is open at the start of the chunk
    for c0 in range(0, n, CHUNK):
        d = (text[c0:c0 + CHUNK] >= 48) & (text[c0:c0 + CHUNK] <= 57)
        prev = _shifted_prev(d, carry > 0)
        nxt = torch.empty_like(d)
        nxt[:-1] = d[1:]
        nxt[-1] = bool(((text[c0 + CHUNK] >= 48) & (text[c0 + CHUNK] <= 57)).item()) if c0 + CHUNK < n else False
        starts = torch.nonzero(d & ~prev).flatten()
        ends = torch.nonzero(d & ~nxt).flatten() + 1     # runs that end inside this chunk
        if bool(d[0].item()) and carry > 0:              # the first ending run began in an earlier chunk
            starts = torch.cat([torch.tensor([-carry], device=text.device), starts])
        k = min(starts.numel(), ends.numel())
        n_matches += int(((ends[:k] - starts[:k]) // 10).sum().item())
        if starts.numel() > ends.numel():                # a run stays open across the chunk end
            carry = d.numel() - int(starts[-1].item())
        else:
            carry = 0
        del d, prev, nxt
    cfg = corpora.CONFIGS["c4"]
    ob = mr.compile_regex(cfg["pattern"]).sub_device(cfg["repl"], mr.Text.from_tensor(text))
    out = ob.tensor()
    assert out.numel() == n + 2 * n_matches
    assert int((out == 45).sum().item()) == 2 * n_matches
    # removing the inserted dashes gives the input back, byte for byte
    done = 0
    for c0 in range(0, out.numel(), CHUNK):
        piece = out[c0:c0 + CHUNK]
        kept = piece[piece != 45]
        assert torch.equal(kept, text[done:done + kept.numel()])
        done += kept.numel()
    assert done == n
    # every dash sits inside ddd-ddd-dddd
    dash = torch.nonzero(out[: 1 << 26] == 45).flatten()
    dash = dash[(dash > 8) & (dash < (1 << 26) - 8)]
    first = dash[0::2][:100000]
    for off in (-3, -2, -1, 1, 2, 3, 5, 6, 7, 8):
        c = out[first + off]
        assert bool(((c >= 48) & (c <= 57)).all().item())
    assert bool((out[first + 4] == 45).all().item())
    ob.close()


def test_c5_full_size_split_equals_mask_compaction():
    n = 4 << 30
    text = corpora.make("c5", n, device=_dev())
    pc = mr.compile_regex(r"\s*,\s*").split_device(mr.Text.from_tensor(text))
    offs = pc.offsets_tensor()

    class _V:
        pass

    total = int(offs[-1].item())
    v = _V()
    v.__cuda_array_interface__ = {"shape": (total,), "typestr": "|u1", "data": (pc.bytes_ptr, False), "version": 3,
                                  "strides": None}
    data = torch.as_tensor(v, device="cuda")
    n_commas = 0
    done = 0
    H = 64
    for c0 in range(0, n, CHUNK):
        lo, hi = max(0, c0 - H), min(n, c0 + CHUNK + H)
        w = text[lo:hi]
        comma = w == 44
        ws = (w == 32) | (w == 9) | (w == 10) | (w == 13) | (w == 12)
        # separator bytes = commas and every whitespace byte connected to a comma through whitespace
        sep = comma.clone()
        for it in range(12):
            grown = sep.clone()
            grown[1:] |= sep[:-1]
            grown[:-1] |= sep[1:]
            new = comma | (ws & grown)
            if it >= 8:
                assert torch.equal(new, sep)     # whitespace runs are short: the dilation has converged
            sep = new
        a, b = c0 - lo, c0 - lo + min(CHUNK, n - c0)
        w, comma, keep = w[a:b], comma[a:b], ~sep[a:b]
        kept = w[keep]
        assert torch.equal(kept, data[done:done + kept.numel()])
        # offsets: piece k + 1 starts at the number of kept bytes before the k-th comma
        before = torch.cumsum(keep.to(torch.int32), 0, dtype=torch.int64) - keep.to(torch.int64) + done
        cpos = torch.nonzero(comma).flatten()
        assert torch.equal(offs[1 + n_commas: 1 + n_commas + cpos.numel()], before[cpos])
        n_commas += cpos.numel()
        done += kept.numel()
        del comma, keep, kept, before, sep, ws, grown, new
    assert done == total
    assert pc.n == n_commas + 1
    assert int(offs[0].item()) == 0
    pc.close()
