"""Parity tests proper: the CUDA path, called through the C ABI (ctypes), against the oracle.

Bit-exact bar: spans, sub output bytes and split pieces must be identical.
"""
import os
import random

import numpy as np
import pytest
import torch

import mojo_regex_b200 as mr
import oracle
from _golden import check_vector, load_vectors
from mojo_regex_b200 import corpora

pytestmark = pytest.mark.gpu

DATA = load_vectors()
TILE = 36864


class _Impl:
    """Adapter so tests/_golden.check_vector can drive the product."""

    class _Rx:
        def __init__(self, p):
            self.rx = mr.compile_regex(p)

        def findall(self, text):
            return self.rx.match_all(text).spans

        def search(self, text):
            m = self.rx.match_next(text, 0)
            return m.span() if m else None

        def match_first(self, text):
            m = self.rx.match_first(text, 0)
            return m.span() if (m and m.start_idx == 0) else None

        def sub(self, repl, text, count=0):
            return self.rx.sub(repl, text, count)

        def split(self, text, maxsplit=0):
            return self.rx.split(text, maxsplit)

    def compile(self, p):
        return self._Rx(p)


@pytest.mark.parametrize("v", DATA["vectors"], ids=lambda v: "%s:%s" % (v["op"], v["src"].split("/")[-1]))
def test_reference_vector_on_gpu(v):
    check_vector(_Impl(), v)


def _compare_all_ops(pattern, text, repl=None):
    orx = oracle.compile(pattern)
    grx = mr.compile_regex(pattern)
    want = orx.findall(text).tolist()
    got = grx.match_all(text).tolist()
    assert got == want, (pattern, text[:200], got[:10], want[:10])
    info = grx.info()
    if info.search_mode != -1:
        for start in (0, 1, len(text) // 2, len(text)):
            w = orx.match_next(text, start)
            m = grx.match_next(text, start)
            assert (m.span() if m else None) == w, (pattern, start, text[:200])
        for start in (0, 3, len(text)):
            w = orx.match_first_at(text, start)
            m = grx.match_first(text, start)
            assert (m.span() if m else None) == w, (pattern, start, text[:200])
        for r in ([b"<>", b""] if repl is None else [repl]):
            try:
                w = orx.sub(r, text)
            except oracle.Unsupported:
                with pytest.raises(mr.UnsupportedPattern):
                    grx.sub(r, text)
                continue
            assert grx.sub(r, text) == w, (pattern, r, text[:200])
            assert grx.sub(r, text, 2) == orx.sub(r, text, 2)
    try:
        w = orx.split(text)
    except oracle.Unsupported:
        with pytest.raises(mr.UndefinedBehaviour):
            grx.split(text)
    else:
        assert grx.split(text) == w
        assert grx.split(text, 3) == orx.split(text, 3)


FUZZ = [
    (r".*", "ab\n"),
    (r"\d+", "0123456789ab xyz"),
    (r"[a-c]*", "abcx "),
    (r"\d{2,4}", "0123456789ab "),
    (r"[^ab]+", "abc d"),
    (r"(hello|world|foo)", "helowrdf "),
    (r"(hello|help|helicopter)", "helpicotr "),
    (r"aa", "ab"),
    (r"aba", "ab"),
    (r"abc", "abc "),
    (r"a|ab|abc", "abc "),
    (r"(a|b)*", "abc"),
    (r"(ab|cd)+", "abcd "),
    (r"(abc)+", "abc "),
    (r"(ab)?", "abc"),
    (r"ab*", "abc"),
    (r"a+", "ab"),
    (r".", "ab\n"),
    (r".+", "ab\n\x80"),
    (r".?", "ab\n"),
    (r"[a-z]+[0-9]+", "abz019 A"),
    (r"[a-z]*[0-9]+", "abz019 A"),
    (r"[a-z]*[0-9]*", "abz019 A"),
    (r"[0-9]+\.?[0-9]*", "0129. a"),
    (r"(\d{3})(\d{3})(\d{4})", "0123456789 a"),
    (r"(\d{2})-(\d{2})", "0123456789- a"),
    (r"[A-Z]{2}-[0-9]{3}", "ABZ-019 "),
    (r"[a-zA-Z0-9._%+-]+@[a-zA-Z0-9.-]+\.[a-z]{2,}", "abAB09._%+-@ "),
    (r"\s*,\s*", "ab, \t\n\x0b"),
    (r"\s*;\s*x", "ab; \tx"),
    (r"[a-z]+@[a-z]+", "ab@ "),
    (r"(?:00|33|44)[0-9]{2}[a-z]+", "0345ab "),
    (r"^ab", "ab "),
    (r"ab$", "ab "),
    (r"^", "ab"),
    (r"$", "ab"),
    (r"^$", "ab"),
    (r"", "ab"),
    (r"\w+\s\w+", "ab_9 \t"),
    (r"[+]*\d+[-]*\d+", "+-019 "),
    # the reference's LazyDFA route (NFAMatcher, matcher.mojo:273-431): tables from the eager subset construction
    (r"\s+", "ab \t\n\x0b"),
    (r"(\d+)", "019a "),
    (r"\d|\w", "a9_ @"),
    (r"(\w\d)+", "a1b2 _9x"),
    (r"(a|b)x", "abx "),
    (r"\w*@", "ab_@ "),
    (r"3[02]|40|[68]9", "0234689 "),
    (r"[^ab]+x", "abx^c "),
    (r"(a|ab)(c|bcd)", "abcd "),
    (r"(ab)*c?", "abc "),
    (r"(?:ab) (\w+)", "ab _9"),
    (r"(x|y)*z{2,3}", "xyz "),
    # match_next through the MemchrPrefilter (matcher.mojo:784-796): literal prefix / literal inside / at the end
    (r"hello[a-z][A-Z]{2}", "heloAB "),
    (r"[A-Z]{2}@x[a-z]", "AB@xa "),
    (r"[a-z]hello[a-z]", "helo "),
    (r"[a-z][A-Z]{2}ab", "abAB "),
]


@pytest.mark.parametrize("pattern,alphabet", FUZZ, ids=[p or "<empty>" for p, _ in FUZZ])
def test_fuzz_small_texts_against_oracle(pattern, alphabet):
    rng = random.Random(hash(pattern) & 0xFFFFF)
    sizes = [0, 1, 2, 3, 7, 15, 16, 17, 31, 33, 64, 100, 143, 144, 145, 300, 1000]
    for n in sizes + [rng.randrange(0, 400) for _ in range(12)]:
        text = bytes(ord(rng.choice(alphabet)) for _ in range(n))
        _compare_all_ops(pattern, text)


@pytest.mark.parametrize("pattern,alphabet", [(p, a) for p, a in FUZZ if p in (
    r"\d+", r"(hello|world|foo)", r"aa", r"[a-z]+[0-9]+", r"(\d{3})(\d{3})(\d{4})", r"\s*,\s*", r"\d{2,4}",
    r"[a-zA-Z0-9._%+-]+@[a-zA-Z0-9.-]+\.[a-z]{2,}", r"a|ab|abc", r"(ab|cd)+", r"[a-c]*")],
    ids=lambda x: x if isinstance(x, str) and len(x) > 3 else None)
def test_fuzz_tile_boundaries_against_oracle(pattern, alphabet):
    rng = random.Random(hash(pattern) & 0xFFFF)
    for n in (TILE - 1, TILE, TILE + 1, 2 * TILE + 17, 5 * TILE - 3):
        text = bytes(ord(rng.choice(alphabet)) for _ in range(n))
        _compare_all_ops(pattern, text)


@pytest.mark.parametrize("kind", ["all_digits", "no_digits", "straddle"])
def test_c1_edge_variants(kind):
    n = 3 * TILE + 100
    if kind == "all_digits":
        text = b"7" * n
    elif kind == "no_digits":
        text = b"x" * n
    else:
        text = bytearray(b"x" * n)
        for k in range(1, 4):
            text[k * TILE - 3:k * TILE + 4] = b"1234567"
        text = bytes(text)
    for pat in (r"\d+", r"(\d{3})(\d{3})(\d{4})", r"\d{2,4}", r"\d"):
        _compare_all_ops(pat, text, repl=rb"[\1]" if "(" in pat else b"#")


def _corpus(config, size):
    return corpora.make(config, size).numpy()


def test_config1_findall_digits():
    text = _corpus("c1", 4 << 20)
    want = oracle.compile(r"\d+").findall(text)
    got = mr.compile_regex(r"\d+").match_all(text).spans
    assert np.array_equal(got, want)
    assert len(got) > 200_000


def test_config2_findall_alternation():
    text = _corpus("c2", 8 << 20)
    want = oracle.compile(r"(hello|world|foo)").findall(text)
    got = mr.compile_regex(r"(hello|world|foo)").match_all(text).spans
    assert np.array_equal(got, want)
    assert len(got) > 100_000


def test_config3_batch_search_email():
    pat = corpora.CONFIGS["c3"]["pattern"]
    n_lines = 200_000
    text = _corpus("c3", n_lines)
    orx = oracle.compile(pat)
    want = np.full((n_lines, 2), -1, dtype=np.int64)
    for i in range(n_lines):
        r = orx.search(text[i * 128:(i + 1) * 128])
        if r:
            want[i] = r
    grx = mr.compile_regex(pat)
    got = grx.search_batch(text, doc_stride=128)
    assert np.array_equal(got, want)
    assert (want[:, 0] >= 0).sum() > 10_000
    # ragged documents through explicit offsets
    rng = np.random.default_rng(3)
    cuts = np.unique(np.concatenate([[0, text.size], rng.integers(0, text.size, 5000)]))
    got2 = grx.search_batch(text, doc_offsets=cuts)
    for k in rng.integers(0, cuts.size - 1, 400):
        r = orx.search(text[cuts[k]:cuts[k + 1]])
        assert tuple(got2[k]) == (r if r else (-1, -1))


def test_config4_sub_phone_numbers():
    cfg = corpora.CONFIGS["c4"]
    text = _corpus("c4", 4 << 20)
    want = oracle.compile(cfg["pattern"]).sub(cfg["repl"].encode(), text)
    got = mr.compile_regex(cfg["pattern"]).sub(cfg["repl"], text)
    assert got == want
    assert len(got) > len(text)


def test_config5_split_csv():
    text = _corpus("c5", 4 << 20)
    wd, wo = oracle.compile(r"\s*,\s*").split_raw(text)
    gd, go = mr.compile_regex(r"\s*,\s*").split_raw(text)
    assert np.array_equal(go, wo)
    assert gd == wd
    assert len(go) > 400_000


def test_device_resident_text_and_results():
    text = corpora.make("c2", 2 << 20)
    dev = text.cuda()
    t = mr.Text.from_tensor(dev)
    rx = mr.compile_regex(r"(hello|world|foo)")
    ds = rx.match_all_device(t)
    want = oracle.compile(r"(hello|world|foo)").findall(text.numpy())
    assert ds.n == len(want)
    assert np.array_equal(ds.to_numpy(), want)
    ds.close()
    # unaligned device pointer -> non-TMA staging path
    t2 = mr.Text.from_tensor(dev[5:])
    want2 = oracle.compile(r"(hello|world|foo)").findall(text.numpy()[5:])
    assert np.array_equal(rx.match_all(t2).spans, want2)


def test_windowed_text_reports_only_core_keys():
    text = corpora.make("c2", 1 << 20).numpy()
    rx = mr.compile_regex(r"(hello|world|foo)")
    want = oracle.compile(r"(hello|world|foo)").findall(text)
    t = mr.Text(text)
    lo, hi = 300_001, 700_003
    t.set_window(lo, hi, 1000)
    got = rx.match_all(t).spans
    sel = want[(want[:, 0] >= lo) & (want[:, 0] < hi)] + 1000
    assert np.array_equal(got, sel)


@pytest.mark.parametrize("config,size", [("c2", 400_000), ("c1", 250_000)])
def test_sharded_findall_emulated_ranks(config, size):
    """The multi-GPU path with the ranks run one after another on this GPU: every rank scans its
    windowed shard through the C ABI, records are merged exactly as the NCCL all-gather would."""
    from mojo_regex_b200 import distributed as mrd

    world = 3
    pat = corpora.CONFIGS[config]["pattern"]
    rx = mr.compile_regex(pat)
    class_runs = config == "c1"
    shards = [mrd.make_shard(config, size, r, world, "cuda") for r in range(world)]
    locals_ = [rx.match_all(sh.text).spans for sh in shards]
    records = [mrd.findall_record(locals_[r], shards[r]) for r in range(world)]
    got = np.concatenate([mrd.finish_findall(locals_[r], records, r, class_runs, bool(rx.info().findall_direct))[0] for r in range(world)])
    whole = torch.cat([mrd._core(config, size, r, "cuda") for r in range(world)]).cpu().numpy()
    want = oracle.compile(pat).findall(whole)
    assert np.array_equal(got, want)


# ---- host-buffer entry point (mrx_findall_host): chunked upload / scan / download pipeline --------------
PIPE_CHUNK = 32 << 20


@pytest.mark.parametrize("config", ["c2", "c1"])
def test_findall_host_pipelined_matches_oracle(config):
    """Texts of several pipeline chunks: the spans streamed back chunk by chunk must equal the oracle's, and
    the device-resident path's, bit for bit (dense literal set and class runs are the pipelined plans)."""
    pat = corpora.CONFIGS[config]["pattern"]
    n = 3 * PIPE_CHUNK + 12345
    text = corpora.make(config, n)
    if config == "c1":
        # a digit run across every chunk boundary, and one that covers a whole tile
        for b in (PIPE_CHUNK, 2 * PIPE_CHUNK, 3 * PIPE_CHUNK):
            text[b - 7:b + 9] = ord("7")
        text[PIPE_CHUNK + 5 * TILE - 100:PIPE_CHUNK + 6 * TILE + 100] = ord("3")
    else:
        lit = torch.tensor(list(b"hello"), dtype=torch.uint8)
        for b in (PIPE_CHUNK, 2 * PIPE_CHUNK):
            text[b - 2:b + 3] = lit          # a literal straddling the chunk boundary
            text[b - 40:b - 35] = lit
    host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    host.copy_(text)
    rx = mr.compile_regex(pat)
    want = oracle.compile(pat).findall(text.numpy())
    got = rx.match_all_host(host.data_ptr(), n)
    assert np.array_equal(got.spans, want)
    got2 = rx.match_all_host(host.data_ptr(), n)      # second call: capacity from the remembered density
    assert np.array_equal(got2.spans, want)
    # pageable source memory goes through the same entry point
    got3 = rx.match_all(text.numpy())
    assert np.array_equal(got3.spans, want)
    # windowed shard of the same buffer
    lo, hi = PIPE_CHUNK - 3, 2 * PIPE_CHUNK + 11
    got4 = rx.match_all_host(host.data_ptr(), n, lo, hi, 5000).spans
    if config == "c1":   # class runs are clipped to the window (stitched by the caller)
        sel = want[(want[:, 1] > lo) & (want[:, 0] < hi)].copy()
        sel[:, 0] = np.maximum(sel[:, 0], lo)
        sel[:, 1] = np.minimum(sel[:, 1], hi)
    else:
        sel = want[(want[:, 0] >= lo) & (want[:, 0] < hi)]
    assert np.array_equal(got4, sel + 5000)


def test_findall_host_small_and_unpipelined_plans():
    rx = mr.compile_regex(r"[a-z]+[0-9]+")           # attempt plan: uploaded, then scanned
    text = np.frombuffer(b"hello123 world456 test789 " * 1000, dtype=np.uint8)
    want = oracle.compile(r"[a-z]+[0-9]+").findall(text)
    assert np.array_equal(rx.match_all(text).spans, want)
    empty = np.zeros(0, dtype=np.uint8)
    assert len(rx.match_all(empty)) == 0
    assert len(mr.compile_regex(r"(hello|world|foo)").match_all(empty)) == 0


# ---- batch search through the reverse-automaton kernel (k_lines) -------------------------------------
BATCH_PATTERNS = [
    (corpora.CONFIGS["c3"]["pattern"], "ab.@-_%+Z9 \n"),
    (r"[a-z]+[0-9]+", "ab1 2-"),
    (r"[A-Z]{2}-[0-9]{3}", "AZ-019 "),
    (r"(hello|world|foo)", "helowrdf "),
    (r"(\d{3})(\d{3})(\d{4})", "0123456789 a"),
    (r"[89]00[0-9]+", "8900 1a"),
    (r"a+b", "ab c"),
]


@pytest.mark.parametrize("pattern,alphabet", BATCH_PATTERNS, ids=[p for p, _ in BATCH_PATTERNS])
def test_search_batch_lines_kernel_against_oracle(pattern, alphabet):
    import os

    rng = random.Random(hash(pattern) & 0xFFFF)
    rx = mr.compile_regex(pattern)
    orx = oracle.compile(pattern)
    assert rx.info().search_rev_states > 0
    alpha = alphabet.encode()
    # fixed-size documents (vector path): 64-byte documents, 3000 of them
    text = np.frombuffer(bytes(rng.choice(alpha) for _ in range(64 * 3000)), dtype=np.uint8)
    got = rx.search_batch(text, doc_stride=64)
    for i in range(3000):
        r = orx.search(text[64 * i:64 * i + 64])
        assert tuple(got[i]) == (r if r else (-1, -1)), (pattern, i, bytes(text[64 * i:64 * i + 64]))
    # ragged documents, some of them empty, last one short; odd stride -> byte path
    cuts = sorted(set([0, text.size] + [rng.randrange(text.size) for _ in range(2000)]))
    cuts = np.array(cuts + [text.size], dtype=np.int64)  # repeated last offset = empty document
    got2 = rx.search_batch(text, doc_offsets=cuts)
    for k in range(len(cuts) - 1):
        r = orx.search(text[cuts[k]:cuts[k + 1]])
        assert tuple(got2[k]) == (r if r else (-1, -1))
    got3 = rx.search_batch(text, doc_stride=37)
    for i in range(0, text.size // 37, 7):
        r = orx.search(text[37 * i:37 * i + 37])
        assert tuple(got3[i]) == (r if r else (-1, -1))
    # the candidate-queue kernel (k_scan) must agree with the reverse-automaton kernel
    os.environ["MRX_NO_LINES"] = "1"
    try:
        assert np.array_equal(rx.search_batch(text, doc_stride=64), got)
    finally:
        del os.environ["MRX_NO_LINES"]


SKIP_BYTE_PATTERNS = [
    (corpora.CONFIGS["c3"]["pattern"], "@", "ab.-_%+Z9 \n"),
    (r"[A-Z]{2}-[0-9]{3}", "-", "AZ019 "),
    (r"[89]00[0-9]+", "0", "891a "),
]


@pytest.mark.parametrize("stride", [16, 48, 64, 128])
@pytest.mark.parametrize("pattern,skip,alphabet", SKIP_BYTE_PATTERNS, ids=[p for p, _, _ in SKIP_BYTE_PATTERNS])
def test_search_batch_skip_byte_kernel(pattern, skip, alphabet, stride):
    """k_lines_req: documents without the automaton's required byte are answered without a walk, the others
    are compacted into full warps.  Sparse and dense candidate mixes, document counts that are not multiples of
    32, enough documents for the 64-row ring to wrap many times; checked against the oracle document by document
    and against the unfiltered kernel."""
    import os

    rx = mr.compile_regex(pattern)
    assert rx.info().search_skip_byte == ord(skip)
    orx = oracle.compile(pattern)
    rng = np.random.default_rng(stride * 131 + len(pattern))
    for n_docs, p_cand in ((1, 1.0), (31, 0.5), (33, 0.0), (5003, 0.35), (2049, 0.97), (4000, 0.02)):
        without = np.frombuffer(alphabet.encode(), dtype=np.uint8)
        with_ = np.frombuffer((alphabet + skip * 2).encode(), dtype=np.uint8)
        docs = without[rng.integers(0, without.size, (n_docs, stride))]
        cand = rng.random(n_docs) < p_cand
        docs[cand] = with_[rng.integers(0, with_.size, (int(cand.sum()), stride))]
        text = np.ascontiguousarray(docs.reshape(-1))
        got = rx.search_batch(text, doc_stride=stride)
        step = 1 if n_docs <= 2049 else 3
        for i in range(0, n_docs, step):
            r = orx.search(text[stride * i:stride * (i + 1)])
            assert tuple(got[i]) == (r if r else (-1, -1)), (pattern, stride, n_docs, i, bytes(docs[i]))
        os.environ["MRX_LINES_FILTER"] = "0"
        try:
            assert np.array_equal(rx.search_batch(text, doc_stride=stride), got)
        finally:
            del os.environ["MRX_LINES_FILTER"]


# ---- assembly kernels (k_seg_lengths / k_assemble) on shapes that stress the tile logic ---------------------
def _rand_text(alphabet: bytes, n: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), n)].copy()


ASSEMBLE_CASES = [
    # (pattern, replacement, alphabet, bytes)        what it stresses
    (r"a", b"", b"ab", 300_000),                     # half of the segments have no output bytes, output shrinks
    (r"a", b"", b"a", 70_000),                       # empty output
    (r"a", b"X" * 300, b"abbbbbbb", 60_000),         # replacement longer than the staged template -> generic path
    (r"[0-9]+", b"<>", b"0123456789 abc", 500_000),  # class runs
    (r"a*", b"-", b"ab", 120_000),                   # zero-width matches: the extra byte after the replacement
    (r"(\d{3})(\d{4})", rb"\2.\1", b"0123456789 x", 400_000),   # fixed-width groups, reordered
    (r"hello", b"HELLO WORLD", b"helo ", 600_000),   # sparse matches, long gaps (segments spanning many tiles)
    (r"b", b"bb", b"b", 40_000),                     # every byte a match: more segments per tile than records staged
]


@pytest.mark.parametrize("pattern,repl,alphabet,n", ASSEMBLE_CASES, ids=[c[0] + "->" + c[1][:6].decode() for c in ASSEMBLE_CASES])
def test_sub_assembly_shapes_against_oracle(pattern, repl, alphabet, n):
    text = _rand_text(alphabet, n, len(pattern) * 7919 + n)
    want = oracle.compile(pattern).sub(repl, text)
    got = mr.compile_regex(pattern).sub(repl, text)
    assert len(got) == len(want)
    assert got == want
    w2 = oracle.compile(pattern).sub(repl, text, 1000)
    assert mr.compile_regex(pattern).sub(repl, text, 1000) == w2


@pytest.mark.parametrize("pattern,alphabet,n", [(r",", b"a,", 200_000), (r"[;,]", b"abc;,", 333_333), (r"\s*,\s*", b"ab, ", 0),
                                                (r",", b",", 50_000), (r"x", b"abc", 100_000)])
def test_split_assembly_shapes_against_oracle(pattern, alphabet, n):
    if n == 0:   # blanks only next to commas (the reference is undefined on `, ,`): build "ab , ab,ab ,  ab ..."
        rng = np.random.default_rng(5)
        parts = []
        for _ in range(40_000):
            parts.append(b"ab"[: rng.integers(1, 3)])
            parts.append(b" " * rng.integers(0, 3) + b"," + b" " * rng.integers(0, 3))
        text = np.frombuffer(b"".join(parts) + b"ab", dtype=np.uint8)
    else:
        text = _rand_text(alphabet, n, n)
    wd, wo = oracle.compile(pattern).split_raw(text)
    gd, go = mr.compile_regex(pattern).split_raw(text)
    assert np.array_equal(go, wo)
    assert gd == wd
    wd2, wo2 = oracle.compile(pattern).split_raw(text, 77)
    gd2, go2 = mr.compile_regex(pattern).split_raw(text, 77)
    assert np.array_equal(go2, wo2) and gd2 == wd2


# ---- k_dense corner cases: token overflow (recompute path), long literals, many literals -------------------
def test_dense_overflow_and_long_literals_against_oracle():
    # more hits per 144-byte chunk than a thread buffers -> recompute path
    t1 = np.frombuffer(b"foo" * 70_000 + b"fo" + b"foo" * 1000, dtype=np.uint8)
    for pat, dense in ((r"foo", True), (r"(foo|bar)", True), (r"(foo|oof)", False)):  # foo/oof can overlap: k_scan + selection
        rx = mr.compile_regex(pat)
        assert bool(rx.info().findall_dense) == dense
        assert np.array_equal(rx.match_all(t1).spans, oracle.compile(pat).findall(t1))
    # the longest literal the DFA path accepts (20 elements; longer ones classify MEDIUM, SURVEY B-6):
    # complete and truncated copies at every alignment relative to the 80/64-byte chains
    lit = bytes(range(97, 117))                # 20 bytes
    rng = np.random.default_rng(9)
    parts = []
    for _ in range(6000):
        parts.append(bytes(rng.integers(97, 123, rng.integers(0, 120)).astype(np.uint8)))
        parts.append(lit[: rng.integers(15, 21)])
    t2 = np.frombuffer(b"".join(parts), dtype=np.uint8)
    pat2 = lit.decode()
    rx2 = mr.compile_regex(pat2)
    assert rx2.info().findall_dense
    want2 = oracle.compile(pat2).findall(t2)
    assert len(want2) > 100
    assert np.array_equal(rx2.match_all(t2).spans, want2)
    m = rx2.match_next(t2, 12345)
    assert (m.span() if m else None) == oracle.compile(pat2).match_next(t2, 12345)
    # 24 literals of mixed lengths
    words = ["alpha", "beta", "gamma", "delta", "epsilon", "zeta", "eta", "theta", "iota", "kappa", "lambda", "mu",
             "nu", "xi", "omicron", "pi", "rho", "sigma", "tau", "upsilon", "phi", "chi", "psi", "omega"]
    pat3 = "(" + "|".join(words) + ")"
    t3 = _rand_text(b"abcdefghiklmnoprstuxyz ", 400_000, 3)
    inject = np.frombuffer(" ".join(words * 50).encode(), dtype=np.uint8)
    t3[1000:1000 + inject.size] = inject
    try:
        rx3 = mr.compile_regex(pat3)
    except mr.UnsupportedPattern:
        with pytest.raises(oracle.Unsupported):
            oracle.compile(pat3)
        return
    assert np.array_equal(rx3.match_all(t3).spans, oracle.compile(pat3).findall(t3))


# ---- host-buffer sub / split (mrx_sub_host / mrx_split_host): chunked three-stream pipeline ----------------
EDIT_CHUNK = 128 << 20


def test_sub_host_pipelined_matches_oracle():
    cfg = corpora.CONFIGS["c4"]
    n = 3 * EDIT_CHUNK + 4321
    text = corpora.make("c4", n)
    for b in (EDIT_CHUNK, 2 * EDIT_CHUNK):      # a long digit run right where a cut would nominally fall
        text[b - 15:b + 25] = torch.tensor(list(b"1234567890" * 4), dtype=torch.uint8)
    host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    host.copy_(text)
    rx = mr.compile_regex(cfg["pattern"])
    want = oracle.compile(cfg["pattern"]).sub(cfg["repl"].encode(), text.numpy())
    hb = rx.sub_host(cfg["repl"], host.data_ptr(), n)
    assert hb.n == len(want)
    assert hb.tobytes() == want
    hb.close()
    # a replacement limit takes the plain path through the same entry point
    w1 = oracle.compile(cfg["pattern"]).sub(cfg["repl"].encode(), text.numpy()[: 1 << 20], 5)
    h1 = rx.sub_host(cfg["repl"], host.data_ptr(), 1 << 20, 5)
    assert h1.tobytes() == w1
    # no-group replacement, class-run pattern, shrinking output
    w2 = oracle.compile(r"[0-9]+").sub(b"#", text.numpy())
    h2 = mr.compile_regex(r"[0-9]+").sub_host(b"#", host.data_ptr(), n)
    assert h2.tobytes() == w2


def test_split_host_pipelined_matches_oracle():
    n = 3 * EDIT_CHUNK + 999
    text = corpora.make("c5", n)
    host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    host.copy_(text)
    rx = mr.compile_regex(r"\s*,\s*")
    wd, wo = oracle.compile(r"\s*,\s*").split_raw(text.numpy())
    hp = rx.split_host(host.data_ptr(), n)
    assert hp.n == len(wo) - 1
    assert np.array_equal(hp.offsets, wo)
    assert hp.data.tobytes() == wd
    hp.close()
    # maxsplit and small texts take the plain path
    wd2, wo2 = oracle.compile(r"\s*,\s*").split_raw(text.numpy()[:300_000], 7)
    h2 = rx.split_host(host.data_ptr(), 300_000, 7)
    assert np.array_equal(h2.offsets, wo2) and h2.data.tobytes() == wd2
    # a pattern without "dead" bytes (every byte can be part of a match) is not chunked
    t3 = np.frombuffer(b"ab" * 1000, dtype=np.uint8)
    assert mr.compile_regex(r"a").split_raw(t3)[0] == oracle.compile(r"a").split_raw(t3)[0]


def test_findall_host_capacity_overflow_falls_back():
    """More matches than the first capacity guess (one per 8 bytes): the pipelined call notices, finishes the
    upload and repeats the list job with the exact count."""
    mr.clear_regex_cache()      # a fresh regex: no remembered density
    n = 2 * PIPE_CHUNK + 77
    host = torch.full((n,), ord("a"), dtype=torch.uint8).pin_memory()
    host[5::97] = ord("b")
    rx = mr.compile_regex(r"a")
    got = rx.match_all_host(host.data_ptr(), n).spans
    t = host.numpy()
    starts = np.flatnonzero(t == ord("a"))
    assert got.shape[0] == starts.size
    assert np.array_equal(got[:, 0], starts) and np.array_equal(got[:, 1], starts + 1)
    got2 = rx.match_all_host(host.data_ptr(), n).spans      # now sized from the remembered density: pipelined
    assert np.array_equal(got2, got)


@pytest.mark.parametrize("pattern,alphabet", FUZZ, ids=[p or "<empty>" for p, _ in FUZZ])
def test_fuzz_medium_texts_all_ops_against_oracle(pattern, alphabet):
    """Every pattern of the fuzz table on ~200 KB (several scan tiles, dozens of assembly tiles, hundreds of
    segment tiles): findall, search, match_first, sub, split -- and batch search through whichever kernel
    the planner picks -- against the oracle."""
    import zlib

    rng = random.Random(zlib.crc32(pattern.encode()))
    n = 200_000 + rng.randrange(0, 5000)
    text = bytes(ord(rng.choice(alphabet)) for _ in range(n))
    _compare_all_ops(pattern, text)
    rx = mr.compile_regex(pattern)
    info = rx.info()
    if info.search_mode != -1:   # batch search (also `^`-anchored plans: one attempt per document)
        orx = oracle.compile(pattern)
        arr = np.frombuffer(text, dtype=np.uint8)
        for stride in (48, 100):
            got = rx.search_batch(arr, doc_stride=stride)
            for i in range(0, len(got), 17):
                r = orx.search(arr[i * stride:(i + 1) * stride])
                assert tuple(got[i]) == (r if r else (-1, -1)), (pattern, stride, i)


def test_search_batch_wildcard_and_anchored():
    text = np.frombuffer(b"abc xyz|ab||xabc ab".replace(b"|", bytes([10])), dtype=np.uint8)   # 19 bytes, | = newline
    cuts = np.array([0, 8, 11, 12, 17, 19], dtype=np.int64)
    for pat in (r".*", r"^ab", r"^abc", r"^", r"ab$"):
        rx, orx = mr.compile_regex(pat), oracle.compile(pat)
        got = rx.search_batch(text, doc_offsets=cuts)
        for k in range(len(cuts) - 1):
            r = orx.search(text[cuts[k]:cuts[k + 1]])
            assert tuple(got[k]) == (r if r else (-1, -1)), (pat, k)
        got = rx.search_batch(text, doc_stride=4)
        for k in range(5):
            r = orx.search(text[4 * k:4 * k + 4])
            assert tuple(got[k]) == (r if r else (-1, -1)), (pat, k)


def test_search_batch_host_pipelined():
    """Three groups of documents through the upload / search / download pipeline: equal to the one-shot device
    path for every document, and to the oracle on a sample; fixed-size and ragged documents."""
    pat = corpora.CONFIGS["c3"]["pattern"]
    n_lines = 2_300_000                      # 294 MB: more than two 128 MiB groups
    text = corpora.make("c3", n_lines)
    host = text.pin_memory()
    rx, orx = mr.compile_regex(pat), oracle.compile(pat)
    out = torch.empty(2 * n_lines, dtype=torch.int64).pin_memory()
    got = rx.search_batch_host(host.data_ptr(), host.numel(), out.numpy(), doc_stride=128)
    want = rx.search_batch(mr.Text(text.numpy()), doc_stride=128)
    assert np.array_equal(got, want)
    rng = np.random.default_rng(2)
    arr = text.numpy()
    for i in rng.integers(0, n_lines, 3000):
        r = orx.search(arr[128 * i:128 * i + 128])
        assert tuple(got[i]) == (r if r else (-1, -1))
    # ragged documents (some empty), offsets on the host
    cuts = np.unique(np.concatenate([[0, arr.size], rng.integers(0, arr.size, 1_500_000)])).astype(np.int64)
    cuts = np.concatenate([cuts[:1000], cuts[999:1000], cuts[1000:]])      # one empty document
    out2 = np.empty(2 * (cuts.size - 1), dtype=np.int64)
    got2 = rx.search_batch_host(host.data_ptr(), host.numel(), out2, doc_offsets=cuts)
    want2 = rx.search_batch(mr.Text(arr), doc_offsets=cuts)
    assert np.array_equal(got2, want2)
    for k in rng.integers(0, cuts.size - 1, 2000):
        r = orx.search(arr[cuts[k]:cuts[k + 1]])
        assert tuple(got2[k]) == (r if r else (-1, -1))


def test_sub_split_ignore_a_shards_global_offset():
    """A shard's global offset rebases the spans findall reports; sub and split assemble in buffer
    coordinates and must not see it (a multi-GPU C4/C5 rank > 0 used to feed rebased spans to the assembly)."""
    t4 = corpora.make("c4", 3 << 20).numpy()
    t5 = corpora.make("c5", 3 << 20).numpy()
    c4 = corpora.CONFIGS["c4"]
    for text, kind in ((t4, "sub"), (t5, "split")):
        t = mr.Text(text)
        t.set_window(0, text.size, 7 * text.size)          # whole buffer is the core, far-away global offset
        if kind == "sub":
            want = oracle.compile(c4["pattern"]).sub(c4["repl"].encode(), text)
            assert mr.compile_regex(c4["pattern"]).sub(c4["repl"], t) == want
        else:
            wd, wo = oracle.compile(r"\s*,\s*").split_raw(text)
            gd, go = mr.compile_regex(r"\s*,\s*").split_raw(t)
            assert gd == wd and np.array_equal(go, wo)
        # findall on the same text does report global coordinates
        rx = mr.compile_regex(r"\d+" if kind == "sub" else r",")
        assert int(rx.match_all(t).spans[0, 0]) >= 7 * text.size


def _supported_random_patterns():
    from test_cabi_cpu import _random_patterns

    out = []
    for p in _random_patterns(420, 20261018):
        try:
            oracle.frontend.compile_regex(p.encode())
            out.append(p)
        except (oracle.Unsupported, oracle.RegexError):
            pass
    return out


@pytest.mark.parametrize("pattern", _supported_random_patterns(), ids=lambda p: p or "<empty>")
def test_random_patterns_all_ops_against_oracle(pattern):
    """Every DFA-routed pattern of the random grammar (all nine plan modes): all ops on a short text and on
    one that crosses a scan tile, alphabet = the pattern's own characters plus a few that match nothing."""
    import zlib

    info = mr.compile_regex(pattern).info()
    if info.findall_mode == -1 or info.search_mode == -1:
        pytest.skip("documented refusal: the reference's result depends on the host SIMD width (DESIGN.md section 2)")
    rng = random.Random(zlib.crc32(pattern.encode()))
    chars = sorted(set(c for c in pattern if c.isalnum() or c in "-.,@;_%+ ")) or ["a"]
    if "\\d" in pattern or "[0-9]" in pattern or "[89]" in pattern:
        chars += list("0189")
    if "\\w" in pattern or "[a-z]" in pattern or "[a-c]" in pattern or "[a-zA-Z0-9._%+-]" in pattern:
        chars += list("abcz")
    if "[A-Z]" in pattern:
        chars += list("AZ")
    if "\\s" in pattern:
        chars += [" ", "\t"]
    alphabet = "".join(chars) + " #\n"
    for n in (700, TILE + 3000):
        text = bytes(ord(rng.choice(alphabet)) for _ in range(n))
        _compare_all_ops(pattern, text)


# ---- round-2 regressions (advisor findings) ---------------------------------------------------------------
@pytest.mark.parametrize("pattern,lit", [(r"(hello|world|foo)", b"hello"), (r"hello", b"hello")])
@pytest.mark.parametrize("edge", [0, 1, 2, 3, 4])
def test_dense_window_ending_on_a_tile_boundary(pattern, lit, edge):
    """A match that STARTS inside the key window but ENDS past the last tile the window touches must still be
    reported (the dense kernel records a match in the tile that holds its end byte): core_end on, or within
    max_len - 1 of, a tile boundary with a literal straddling it -- device path and pipelined host path."""
    n = 40 * TILE + 777
    text = corpora.make("c2", n)
    lo = 16 * 3                        # origin = lo & ~15 = 48
    cut = 48 + 9 * TILE                # a tile boundary of this window's tiling
    hi = cut - edge                    # core_end = boundary - edge: the literal below starts inside, ends outside
    text[cut - edge - 1:cut - edge - 1 + len(lit)] = torch.tensor(list(lit), dtype=torch.uint8)
    rx = mr.compile_regex(pattern)
    want = oracle.compile(pattern).findall(text.numpy())
    sel = want[(want[:, 0] >= lo) & (want[:, 0] < hi)]
    assert (sel[:, 1] > hi).any()      # the straddling literal belongs to this window
    t = mr.Text(text.numpy())
    t.set_window(lo, hi, 0)
    assert np.array_equal(rx.match_all(t).spans, sel)
    # host path (pipelined from 16 MiB on: use a larger buffer with the same edge geometry)
    n2 = (17 << 20) + 333
    big = corpora.make("c2", n2)
    cut2 = 48 + 300 * TILE
    hi2 = cut2 - edge
    big[hi2 - 1:hi2 - 1 + len(lit)] = torch.tensor(list(lit), dtype=torch.uint8)
    host = torch.empty(n2, dtype=torch.uint8, pin_memory=True)
    host.copy_(big)
    want2 = oracle.compile(pattern).findall(big.numpy())
    sel2 = want2[(want2[:, 0] >= lo) & (want2[:, 0] < hi2)]
    assert (sel2[:, 1] > hi2).any()
    got2 = rx.match_all_host(host.data_ptr(), n2, lo, hi2, 0).spans
    assert np.array_equal(got2, sel2)


@pytest.mark.parametrize("pattern", [r"\d{10}", r"[a-z]{2,4}", r"(\d{3})(\d{3})(\d{4})"])
@pytest.mark.parametrize("lo", [5, 16, 37, TILE - 3, TILE + 16, 2 * TILE + 7])
def test_uniform_run_window_keeps_the_phase_of_a_straddling_run(pattern, lo):
    """UNIFORM_RUN on a windowed text: a class run that straddles core_begin is chunked with the phase it has
    in the whole buffer (the reference sees the whole text); the window only filters by match start."""
    n = 4 * TILE + 100
    text = corpora.make("c4", n).numpy().copy()
    fill = ord("7") if "d" in pattern else ord("q")
    # one long run across `lo`, starting 1..40 bytes before it, and one that swallows a whole tile before it
    text[max(lo - 33, 0):lo + 57] = fill
    if lo > TILE:
        text[3:lo - 40] = fill
        text[lo - 40] = ord(" ")
    rx = mr.compile_regex(pattern)
    assert rx.info().findall_mode == 8
    want = oracle.compile(pattern).findall(text)
    hi = n - 11
    sel = want[(want[:, 0] >= lo) & (want[:, 0] < hi)]
    t = mr.Text(text)
    t.set_window(lo, hi, 123)
    assert np.array_equal(rx.match_all(t).spans, sel + 123)


def test_from_tensor_orders_after_the_producer_stream():
    """Text.from_tensor on a tensor that torch kernels are still writing: the scan must see the final bytes."""
    rx = mr.compile_regex(r"(hello|world|foo)")
    base = corpora.make("c2", 8 << 20).cuda()
    want = oracle.compile(r"(hello|world|foo)").findall(base.cpu().numpy())
    for _ in range(5):
        junk = torch.zeros_like(base)
        buf = torch.cat([junk[:0], base])          # produced on torch's stream right before the scan
        t = mr.Text.from_tensor(buf)
        ds = rx.match_all_device(t)
        assert ds.n == len(want)
        ds.close()


def test_sub_through_a_packed_automaton_with_fixed_width_metadata():
    """The Mojo serializer's route (mrx_dfa_pack -> mrx_regex_load) for BASELINE config 4: raw DFAEngine rows
    plus CompiledRegex._fixed_sub_* (matcher.mojo:1002-1035) -> sub with group references."""
    pat = rb"(\d{3})(\d{3})(\d{4})"
    from _pack import pack_through_the_serializer_entry

    blob = pack_through_the_serializer_entry(pat)
    assert blob == mr.CompiledRegex(pat).blob()
    rx = mr.CompiledRegex(blob=blob)
    i = rx.info()
    assert (i.fixed_sub_total_width, i.fixed_sub_num_groups, i.fixed_sub_concat) == (10, 3, 1)
    eng = oracle.compile(pat)
    text = corpora.make("c4", 1 << 18).numpy()
    assert rx.sub(r"\1-\2-\3", text) == eng.sub(rb"\1-\2-\3", text)
    assert rx.sub(r"\1-\2-\3", b"6502530000 and 4155551234", 1) == b"650-253-0000 and 4155551234"
    # without the metadata (what the round-1 stub packed) group references are refused, loudly
    rx_bare = mr.CompiledRegex(blob=mr.CompiledRegex(rb"\d{10}").blob())   # same automaton shape, no groups
    with pytest.raises(mr.UnsupportedPattern):
        rx_bare.sub(r"\1-\2-\3", b"6502530000")


# ---- k_sand (Shift-And literal-set scan) against the oracle and against k_dense ----------------------------------
SAND_CASES = [
    (r"(hello|world|foo)", "c2"), (r"hello", "c2"), (r"(foo|he)", "c2"), (r"a", "c2"), (r"(ab|cd|ef|gh)", "c2"),
    (r"ing ", "c2"), (r"(hell|worl|fo)", "c2"), (r"aa", "c2"),
]


def _fresh_regex(pattern, tail, monkeypatch):
    """Compiled outside the pattern cache: MRX_SAND_NO_TAIL is read when the plan is made (tail = the form of the
    Shift-And automaton whose hit bits are read once per text word, sand.hpp)."""
    if tail:
        monkeypatch.delenv("MRX_SAND_NO_TAIL", raising=False)
    else:
        monkeypatch.setenv("MRX_SAND_NO_TAIL", "1")
    return mr.CompiledRegex(pattern)


@pytest.mark.parametrize("pattern,config", SAND_CASES)
@pytest.mark.parametrize("tail", [True, False])
def test_sand_kernel_equals_oracle_and_dense_kernel(pattern, config, tail, monkeypatch):
    rx = _fresh_regex(pattern, tail, monkeypatch)
    assert rx.info().findall_dense
    orx = oracle.compile(pattern)
    for n in (0, 1, 15, 143, 144, 145, TILE - 1, TILE, TILE + 1, 3 * TILE + 77, 2_000_003):
        text = corpora.make(config, max(n, 1))[:n].numpy()
        want = orx.findall(text)
        got = rx.match_all(mr.Text(text)).spans
        assert np.array_equal(got, want), (pattern, n, tail)
        monkeypatch.setenv("MRX_NO_SAND", "1")
        dense = rx.match_all(mr.Text(text)).spans
        monkeypatch.delenv("MRX_NO_SAND")
        assert np.array_equal(dense, want), (pattern, n, "k_dense")


@pytest.mark.parametrize("tail", [True, False])
def test_sand_kernel_windows_unaligned_and_dense_hits(tail, monkeypatch):
    rx = _fresh_regex(r"(hello|world|foo)", tail, monkeypatch)
    orx = oracle.compile(r"(hello|world|foo)")
    text = corpora.make("c2", 5 * TILE + 1234)
    # literals across every tile boundary of the origin-0 tiling and of the window tilings below, hit-dense stretches
    # (one hit every 3 bytes: more hits per chunk than any private list held), a literal at the very end
    for b in (TILE, 2 * TILE, 3 * TILE, 48 + TILE, 16 + 2 * TILE):
        for d in (-4, -2, -1, 0):
            text[b + d:b + d + 5] = torch.tensor(list(b"hello"), dtype=torch.uint8)
    text[1000:4000] = torch.tensor(list(b"foo" * 1000), dtype=torch.uint8)
    text[-3:] = torch.tensor(list(b"foo"), dtype=torch.uint8)
    host = text.numpy()
    want = orx.findall(host)
    assert np.array_equal(rx.match_all(mr.Text(host)).spans, want)
    for lo, hi in [(5, len(host)), (16, TILE + 16), (48, 48 + TILE), (TILE - 2, TILE + 3), (1001, 3999), (0, 0), (7, 7)]:
        t = mr.Text(host)
        t.set_window(lo, hi, 10**12)
        sel = want[(want[:, 0] >= lo) & (want[:, 0] < hi)] + 10**12
        assert np.array_equal(rx.match_all(t).spans, sel), (lo, hi)
    # unaligned device pointer: staging without TMA
    dev = text.cuda()
    for off in (1, 5, 17):
        got = rx.match_all(mr.Text.from_tensor(dev[off:])).spans
        assert np.array_equal(got, orx.findall(host[off:])), off
    # text with bytes >= 0x80 and NUL bytes (a literal containing NUL would see stale buffer bytes past the text)
    raw = np.frombuffer(bytes(random.Random(3).choices(range(256), k=200_000)), dtype=np.uint8).copy()
    raw[5000:5005] = np.frombuffer(b"hello", dtype=np.uint8)
    assert np.array_equal(rx.match_all(mr.Text(raw)).spans, orx.findall(raw))


# ---- k_split (fused split for X* c Y* patterns) against the oracle and against the general path ------------------
def _split_both_paths(rx, text, monkeypatch):
    """-> (fused result, general-path result); each is (bytes, offsets list) or the exception class raised."""
    res = []
    for fused in (True, False):
        if fused:
            monkeypatch.delenv("MRX_NO_FUSED_SPLIT", raising=False)
        else:
            monkeypatch.setenv("MRX_NO_FUSED_SPLIT", "1")
        try:
            d, o = rx.split_raw(text)
            res.append((d, o.tolist()))
        except mr.UndefinedBehaviour:
            res.append(mr.UndefinedBehaviour)
    monkeypatch.delenv("MRX_NO_FUSED_SPLIT", raising=False)
    return res


@pytest.mark.parametrize("pattern,alphabet", [(r"\s*,\s*", b"ab,,, \t\n\r\x0b\x0c\xa0\x8c"), (r"\s*;\s*", b"ab; \t\n\x0b"),
                                              (r"[ab]*,[ab]*", b"abc,, "), (r"\d*-\d*", b"019-x- "), (r"\s*=\s*", b"k=v \t=")])
def test_fused_split_random_texts(pattern, alphabet, monkeypatch):
    rx = mr.compile_regex(pattern)
    orx = oracle.compile(pattern)
    rng = random.Random(77)
    outcomes = set()
    for _ in range(70):
        n = rng.choice([1, 2, 31, 95, 96, 143, 144, 145, 300, 1000, TILE - 1, TILE, TILE + 1, 2 * TILE + 55])
        if rng.random() < 0.6:   # separators with their own blanks only: no shared runs
            sep = bytes([rx.info().required_byte])
            pad = bytes(b for b in alphabet if b not in sep and b not in b"abkvxc019" and b < 0x80) or b""
            toks = []
            while sum(map(len, toks)) < n:
                toks.append(bytes(rng.choice(b"abkvxc019") for _ in range(rng.randint(0, 7))) or b"q")
                toks.append(bytes(rng.choice(pad) for _ in range(rng.randint(0, 2))) + sep +
                            bytes(rng.choice(pad) for _ in range(rng.randint(0, 2))) if pad else sep)
            t = b"".join(toks)[:n]
        else:
            t = bytes(rng.choice(alphabet) for _ in range(n))
        try:
            wd, wo = orx.split_raw(t)
            want = (wd, wo.tolist())
        except oracle.Unsupported:
            want = mr.UndefinedBehaviour
        fused, general = _split_both_paths(rx, t, monkeypatch)
        assert fused == want, (pattern, t[:300])
        assert general == want, (pattern, t[:300])
        outcomes.add("undefined" if want is mr.UndefinedBehaviour else "ok")
    assert "ok" in outcomes


def _check_split(rx, orx, text):
    try:
        wd, wo = orx.split_raw(text)
    except oracle.Unsupported:
        with pytest.raises(mr.UndefinedBehaviour):
            rx.split_raw(text)
        return None
    gd, go = rx.split_raw(text)
    assert gd == wd and np.array_equal(go, wo)
    return wo


def test_fused_split_c5_corpus_tiles_and_long_runs(monkeypatch):
    rx = mr.compile_regex(r"\s*,\s*")
    orx = oracle.compile(r"\s*,\s*")
    text = corpora.make("c5", 6 * TILE + 4321).numpy().copy()
    wo = _check_split(rx, orx, text)
    assert wo is not None and len(wo) > 20_000
    # separators across every tile boundary, at the very start and the very end
    ok = 0
    for b in (TILE, 2 * TILE, 3 * TILE):
        for d in (-2, -1, 0, 1):
            t2 = text.copy()
            t2[b + d - 3:b + d + 6] = np.frombuffer(b"wx , yzqq", dtype=np.uint8)
            ok += _check_split(rx, orx, t2) is not None
    assert ok >= 6
    t3 = text.copy()
    t3[:5] = np.frombuffer(b" , ab", dtype=np.uint8)
    t3[-5:] = np.frombuffer(b"ab , ", dtype=np.uint8)
    assert _check_split(rx, orx, t3) is not None
    # blank runs up to and beyond what the kernel's windows measure: short ones fused, long ones through the
    # general path -- the result is the reference's either way
    for run in (15, 16, 31, 32, 46, 47, 48, 64, 200, TILE + 10):
        t4 = text[: 5 * TILE].copy()
        pos = TILE - 20
        t4[pos - 2:pos] = np.frombuffer(b"ab", dtype=np.uint8)
        t4[pos:pos + run] = 32
        t4[pos + run] = ord(",")
        t4[pos + run + 1:pos + 2 * run + 1] = 9
        t4[pos + 2 * run + 1:pos + 2 * run + 4] = np.frombuffer(b"zzz", dtype=np.uint8)
        assert _check_split(rx, orx, t4) is not None, run
    # \v in the run before a comma: the comma is no separator
    t5 = np.frombuffer(b"a \x0b ,b , c\x0b,d", dtype=np.uint8)
    assert rx.split(t5) == orx.split(t5)
    # device-resident, unaligned pointer
    dev = torch.from_numpy(text).cuda()
    pc = rx.split_device(mr.Text.from_tensor(dev[7:]))
    wd, wo = orx.split_raw(text[7:])
    assert pc.n == len(wo) - 1
    assert np.array_equal(pc.offsets_tensor().cpu().numpy(), wo)
    pc.close()
    # a second call reuses the remembered separator density for the size of the offsets array; a text with many
    # more separators than the first one makes the kernel run twice
    dense = np.frombuffer(b"a,b,c,d," * 40_000, dtype=np.uint8)
    assert _check_split(rx, orx, dense) is not None


# ---- multi-device mode behind the C ABI (mrx_init_multi): one call, every GPU of the box --------------------------
def _needs_gpus(n):
    return pytest.mark.skipif(torch.cuda.device_count() < n, reason="needs %d GPUs" % n)


@_needs_gpus(2)
def test_multi_device_single_call_equals_single_device():
    """findall / batch search / sub / split over host buffers: the sharded single call on 2 (and, when present, more)
    GPUs returns bit for bit what one GPU returns -- class runs that straddle or swallow a shard included."""
    n_dev = min(torch.cuda.device_count(), 8)
    MiB = 1 << 20
    texts = {}
    for config in ("c1", "c2", "c4", "c5"):
        t = corpora.make(config, n_dev * 40 * MiB + 12345)
        if config == "c1":   # a digit run across the first shard edge, and one that swallows the whole second shard
            edge = t.numel() // n_dev
            t[edge - 9:edge + 7] = ord("5")
            if n_dev >= 3:
                t[edge * 1 + 100:edge * 2 + 50] = ord("8")
        pinned = torch.empty(t.numel(), dtype=torch.uint8, pin_memory=True)
        pinned.copy_(t)
        texts[config] = pinned
    lines = torch.empty(n_dev * 300_000 * 128, dtype=torch.uint8, pin_memory=True)
    lines.copy_(corpora.make("c3", n_dev * 300_000))

    def run_all():
        res = {}
        for config in ("c1", "c2"):
            rx = mr.CompiledRegex(corpora.CONFIGS[config]["pattern"])
            res[config] = rx.match_all_host(texts[config].data_ptr(), texts[config].numel()).spans.copy()
        rx = mr.CompiledRegex(corpora.CONFIGS["c4"]["pattern"])
        hb = rx.sub_host(corpora.CONFIGS["c4"]["repl"], texts["c4"].data_ptr(), texts["c4"].numel())
        res["c4"] = hb.array.copy()
        hb.close()
        rx = mr.CompiledRegex(corpora.CONFIGS["c5"]["pattern"])
        hp = rx.split_host(texts["c5"].data_ptr(), texts["c5"].numel())
        res["c5"] = (hp.data.copy(), hp.offsets.copy())
        hp.close()
        rx = mr.CompiledRegex(corpora.CONFIGS["c3"]["pattern"])
        out = np.empty(2 * n_dev * 300_000, dtype=np.int64)
        res["c3"] = rx.search_batch_host(lines.data_ptr(), lines.numel(), out, doc_stride=128).copy()
        return res

    mr.clear_regex_cache()
    try:
        mr.init_multi(n_dev)
        assert mr.device_count() == n_dev
        multi = run_all()
    finally:
        mr.shutdown()
        mr.init(0)
    single = run_all()
    for config in ("c1", "c2", "c3", "c4"):
        assert np.array_equal(multi[config], single[config]), config
    assert np.array_equal(multi["c5"][0], single["c5"][0]) and np.array_equal(multi["c5"][1], single["c5"][1])
    # and the single-device results are the oracle's (bounded prefix: the oracle is slow)
    want = oracle.compile(corpora.CONFIGS["c1"]["pattern"]).findall(texts["c1"].numpy()[: 8 * MiB])
    got = single["c1"][single["c1"][:, 1] <= 8 * MiB - 64]
    assert np.array_equal(got, want[: len(got)])


@pytest.mark.parametrize("v", DATA["group_sub"], ids=lambda v: "%s:%s" % (v["pattern"], v["src"].split("/")[-1]))
def test_group_sub_is_refused_until_the_capture_kernel_exists(v):
    rx = mr.compile_regex(v["pattern"])
    with pytest.raises(mr.UnsupportedPattern):
        rx.sub(v["repl"], v["text"])
    assert rx.sub("<>", v["text"]) == oracle.compile(v["pattern"]).sub(b"<>", v["text"].encode())


# ---- k_subst (fused sub for fixed-width chains over one class) against the oracle and against the general path -----
def _sub_both_paths(pattern, repl, text, monkeypatch):
    rx = mr.compile_regex(pattern)
    monkeypatch.delenv("MRX_NO_FUSED_SUB", raising=False)
    fused = rx.sub(repl, text)
    monkeypatch.setenv("MRX_NO_FUSED_SUB", "1")
    general = rx.sub(repl, text)
    monkeypatch.delenv("MRX_NO_FUSED_SUB", raising=False)
    return fused, general


SUBST_CASES = [
    # (pattern, replacement, width)                              what it stresses
    (r"(\d{3})(\d{3})(\d{4})", rb"\1-\2-\3", 10),                # config 4: the text grows by 2 bytes per match
    (r"(\d{3})(\d{3})(\d{4})", rb"\3\1", 10),                    # groups out of order, a group dropped: shrinks
    (r"(\d{3})(\d{3})(\d{4})", rb"", 10),                        # matches deleted
    (r"(\d{3})(\d{3})(\d{4})", rb"(\1) \2-\3 [\2\2\2] " + b"x" * 30, 10),   # 59 bytes per replacement
    (r"\d{4}", rb"<#>", 4),                                      # the narrowest width, no group references
    (r"[a-c]{32}", rb"_", 32),                                   # the widest
    (r"([0-9a-f]{4})([0-9a-f]{3})", rb"<hex>", 7),               # class read from the table
    (r"[ab]{5}", rb"ZZZZZZZ", 5),                                # needle filter
]


@pytest.mark.parametrize("pattern,repl,width", SUBST_CASES)
def test_fused_sub_equals_oracle_and_general_path(pattern, repl, width, monkeypatch):
    assert mr.compile_regex(pattern).info().sub_fused_width == width
    orx = oracle.compile(pattern)
    rng = np.random.default_rng(width * 131 + len(repl))
    cls = {10: b"0123456789", 4: b"0123456789", 32: b"abc", 7: b"0123456789abcdef", 5: b"ab"}[width]
    other = b" xyz-,."
    texts = []
    # runs of every length around the width and its multiples, separated by 1..3 other bytes
    parts = []
    for _ in range(30_000):
        k = int(rng.integers(0, 4 * width + 3))
        parts.append(bytes(cls[i] for i in rng.integers(0, len(cls), k)))
        parts.append(bytes(other[i] for i in rng.integers(0, len(other), int(rng.integers(1, 4)))))
    texts.append(b"".join(parts))
    # sparse matches in long stretches of other bytes; a text that is one run; runs across tile and chunk edges
    texts.append(b"".join(bytes(other[i] for i in rng.integers(0, len(other), int(rng.integers(0, 900)))) +
                          bytes(cls[i] for i in rng.integers(0, len(cls), width)) for _ in range(600)))
    texts.append(bytes(cls[i] for i in rng.integers(0, len(cls), 100_003)))
    edge = bytearray(b"." * (3 * 36864 + 500))
    for at in (36864, 2 * 36864, 36864 + 144 * 7, 3 * 36864 + 400):
        for back in (1, width - 1, width, 2 * width + 3):
            lo = at - back
            edge[lo:lo + 3 * width + 1] = bytes(cls[i] for i in rng.integers(0, len(cls), 3 * width + 1))
            fused, general = _sub_both_paths(pattern, repl, bytes(edge), monkeypatch)
            assert fused == general == orx.sub(repl, np.frombuffer(bytes(edge), np.uint8)), (at, back)
            edge[lo:lo + 3 * width + 1] = b"." * (3 * width + 1)
    for n_cut in (0, 1, width - 1, width, 143, 144, 145, 36863, 36864, 36865):
        texts.append(texts[0][:n_cut])
        texts.append(texts[2][:n_cut])
    for text in texts:
        fused, general = _sub_both_paths(pattern, repl, text, monkeypatch)
        want = orx.sub(repl, np.frombuffer(text, np.uint8)) if text else b""
        assert len(fused) == len(want)
        assert fused == want
        assert general == want


def test_fused_sub_on_the_config4_corpus_and_with_a_limit(monkeypatch):
    cfg = corpora.CONFIGS["c4"]
    text = _corpus("c4", 16 << 20)
    fused, general = _sub_both_paths(cfg["pattern"], cfg["repl"], text, monkeypatch)
    assert fused == general
    small = text[: 1 << 20]
    assert mr.compile_regex(cfg["pattern"]).sub(cfg["repl"], small) == oracle.compile(cfg["pattern"]).sub(cfg["repl"].encode(), small)
    # a replacement limit is the general path's business
    assert mr.compile_regex(cfg["pattern"]).sub(cfg["repl"], small, 5) == oracle.compile(cfg["pattern"]).sub(cfg["repl"].encode(), small, 5)


def test_fused_sub_giant_run_spanning_many_tiles(monkeypatch):
    """One run of digits over 40 tiles: the phase of every chunk comes from the carry pre-pass."""
    n = 40 * 36864 + 77
    text = np.full(n, ord("7"), dtype=np.uint8)
    text[13] = ord("x")                       # the run starts at 14: starts at 14 + 10 k
    text[n - 40] = ord("y")
    fused, general = _sub_both_paths(r"(\d{3})(\d{3})(\d{4})", rb"\1-\2-\3", text, monkeypatch)
    assert fused == general
    body = (n - 40 - 14) // 10                # + one match in the 13 digits before the x, three in the 39 after the y
    assert len(fused) == n + 2 * (body + 4)


# ---- adversarial shapes: texts that are one giant run ---------------------------------------------------------------
def _timed_ms(fn):
    import time

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    return r, (time.perf_counter() - t0) * 1e3


def _record_timing(name, ms, nbytes):
    """Wall time of adversarial cases, kept for DESIGN.md (gpurun_out/ is scratch: copied to profiles/ by hand)."""
    import json

    os.makedirs("gpurun_out", exist_ok=True)
    path = "gpurun_out/adversarial_timings.json"
    rec = json.load(open(path)) if os.path.exists(path) else {}
    rec[name] = {"ms": round(ms, 3), "bytes": nbytes}
    json.dump(rec, open(path, "w"), indent=1)


def _spans_of(ds):
    return ds.tensor().cpu().numpy().reshape(-1, 2)


def test_adversarial_one_run_of_letters_then_a_digit():
    """`[a-z]+[0-9]+` over a...a1: the reference makes ONE attempt (at 0) that walks the text; every other byte is a
    candidate too, and without the leading-run pruning each would walk to the end of the text (n^2 / 2 bytes)."""
    rx = mr.compile_regex(r"[a-z]+[0-9]+")
    assert rx.info().findall_leading_run == 1
    small = np.full(300_000, ord("a"), dtype=np.uint8)
    small[-1] = ord("7")
    assert np.array_equal(rx.match_all(small).spans, oracle.compile(r"[a-z]+[0-9]+").findall(small))
    for n in (1 << 20, 32 << 20):
        t = torch.full((n,), ord("a"), dtype=torch.uint8, device="cuda")
        t[n - 2:] = ord("1")
        txt = mr.Text.from_tensor(t)
        ds, ms = _timed_ms(lambda: rx.match_all_device(txt))
        assert _spans_of(ds).tolist() == [[0, n]]
        _record_timing("letters_then_digit_findall_%dMiB" % (n >> 20), ms, n)
        m, ms = _timed_ms(lambda: rx.match_next(txt, 5))
        assert (m.start_idx, m.end_idx) == (5, n)
        _record_timing("letters_then_digit_search_%dMiB" % (n >> 20), ms, n)
    # no digit at all: the one attempt fails at the end of the text, the pruned ones would have failed too
    t = torch.full((8 << 20,), ord("q"), dtype=torch.uint8, device="cuda")
    txt = mr.Text.from_tensor(t)
    ds, ms = _timed_ms(lambda: rx.match_all_device(txt))
    assert _spans_of(ds).shape[0] == 0
    _record_timing("letters_only_findall_8MiB", ms, 8 << 20)
    assert rx.match_next(txt, 0) is None


def test_adversarial_leading_run_family_on_mixed_runs():
    """Runs of every length against the oracle: the pruned candidates never carry a reported match."""
    rng = np.random.default_rng(77)
    for pattern, alphabet in ((r"[a-z]+[0-9]+", b"abz019 -"), (r"[ab]+c", b"abcc d"), (r"\d+x", b"0189xx y"),
                              (r"[a-z]+\d*", b"ab12 ,"), (r"a+1a?", b"a1 "), (r"a+b", b"ab ")):
        orx, rx = oracle.compile(pattern), mr.compile_regex(pattern)
        for n in (0, 1, 2, 17, 4097, 150_001):
            text = np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), n)].copy()
            assert np.array_equal(rx.match_all(text).spans, orx.findall(text)), (pattern, n)
            t = mr.Text(text)
            assert np.array_equal(_spans_of(rx.match_all_device(t)), orx.findall(text)), (pattern, n)


def test_adversarial_all_digits_and_a_line_without_the_required_byte():
    n = 256 << 20
    t = torch.full((n,), ord("5"), dtype=torch.uint8, device="cuda")
    txt = mr.Text.from_tensor(t)
    rx = mr.compile_regex(r"\d+")
    ds, ms = _timed_ms(lambda: rx.match_all_device(txt))
    assert _spans_of(ds).tolist() == [[0, n]]
    _record_timing("all_digits_findall_256MiB", ms, n)
    # C4's pattern over all digits: every 10 bytes a match, the phase carried across 7282 tiles
    c4 = corpora.CONFIGS["c4"]
    rx4 = mr.compile_regex(c4["pattern"])
    ob, ms = _timed_ms(lambda: rx4.sub_device(c4["repl"], txt))
    assert ob.n == n + 2 * (n // 10)
    _record_timing("all_digits_sub_256MiB", ms, n)
    head = ob.tensor()[:24].cpu().numpy().tobytes()
    assert head == b"555-555-5555" * 2
    # e-mail search on a 1 MiB line that has no '@': one attempt (the others are dominated) walks the line
    email = r"[a-zA-Z0-9._%+-]+@[a-zA-Z0-9.-]+\.[a-zA-Z]{2,}"
    rxe = mr.compile_regex(email)
    assert rxe.info().search_dominated == 1
    line = torch.full((1 << 20,), ord("k"), dtype=torch.uint8, device="cuda")
    ltxt = mr.Text.from_tensor(line)
    m, ms = _timed_ms(lambda: rxe.match_next(ltxt, 0))
    assert m is None
    _record_timing("email_search_1MiB_line_without_at", ms, 1 << 20)
    ds, ms = _timed_ms(lambda: rxe.match_all_device(ltxt))
    assert _spans_of(ds).shape[0] == 0


# ---- asynchronous findall: no host round trip, the count stays in HBM -----------------------------------------------
@pytest.mark.parametrize("config", ["c1", "c2", "c4"])
def test_findall_async_equals_findall(config):
    cfg = corpora.CONFIGS[config]
    rx = mr.CompiledRegex(cfg["pattern"])                    # fresh: no density estimate yet
    text = corpora.make(config, 6 << 20)
    t = mr.Text(text.numpy())
    want = rx.match_all_device(t).to_numpy()
    results = [rx.match_all_async(t) for _ in range(3)]      # three in flight, released out of order
    assert all(r.n == -1 for r in results)
    cnt = torch.empty(1, dtype=torch.int64, device="cuda")
    for r in (results[1], results[0], results[2]):
        assert r.count_ptr != 0
        assert r.wait()
        assert r.n == len(want)
        assert np.array_equal(r.to_numpy(), want)
        r.close()
    # released while still pending
    rx.match_all_async(t).close()
    # a text denser than the estimate: the array is too small, wait() says so, and the repeat fits
    dense = np.frombuffer((b"7 " if config != "c2" else b"foo ") * 600_000, dtype=np.uint8)
    if config != "c4":
        td = mr.Text(dense)
        r = rx.match_all_async(td)
        ok = r.wait()
        r.close()
        r2 = rx.match_all_async(td)
        assert r2.wait()
        assert r2.n == len(rx.match_all_device(td).to_numpy()) and (ok or r2.n > 0)
        r2.close()


def test_findall_async_on_plans_that_need_the_host():
    """Non-direct plans and empty texts run the ordinary path and park their count on the device."""
    for pattern, text in ((r"a+1a?", b"a1aa1 aa1a"), (r"\d+", b""), (r".*", b"abc"), (r"^ab", b"abab")):
        rx = mr.compile_regex(pattern)
        t = mr.Text(np.frombuffer(text, dtype=np.uint8)) if text else mr.Text(np.zeros(0, np.uint8))
        r = rx.match_all_async(t)
        assert r.wait()
        assert np.array_equal(r.to_numpy().reshape(-1, 2), oracle.compile(pattern).findall(np.frombuffer(text, np.uint8)).reshape(-1, 2))
        r.close()
