"""CPU-side checks of the C ABI: the library loads, exports every symbol the header declares,
compiles patterns without a GPU, refuses to match without one, and its front end builds the
same automata as the oracle's independent restatement."""
import ctypes
import os
import re
import struct

import numpy as np
import pytest

import mojo_regex_b200 as mr
import oracle
from _golden import load_vectors

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "mojoregex_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mrx_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(mr.lib, s), s
    assert sorted(mr.EXPORTS) == syms
    assert mr.lib.mrx_abi_version() == 1


def decode_blob(blob: bytes):
    hdr = struct.unpack_from("<4I4i3i10i10i4I", blob, 0)
    magic, version, total, flags, ns, nc, start, req = hdr[:8]
    assert magic == 0x4258524D and version == 1 and total == len(blob)
    ftw, fng, fcc = hdr[8:11]
    foff, fwid = hdr[11:21], hdr[21:31]
    llen, elen, plen, _ = hdr[31:35]
    off = struct.calcsize("<4I4i3i10i10i4I")
    cls = np.frombuffer(blob, np.uint8, 256, off); off += 256
    lookup = np.frombuffer(blob, np.uint8, 256, off); off += 256
    acc = np.frombuffer(blob, np.uint8, ns, off); off += ns
    tr = np.frombuffer(blob, np.uint8, ns * nc, off).reshape(ns, nc) if ns else np.zeros((0, 0), np.uint8)
    off += ns * nc
    lit = blob[off:off + llen]; off += llen
    exact = blob[off:off + elen]; off += elen
    pref = blob[off:off + plen]
    full = np.full((ns, 256), -1, dtype=np.int32)
    for s in range(ns):
        row = tr[s][cls].astype(np.int32)
        row[row == 255] = -1
        full[s] = row
    return dict(flags=flags, start=start, req=req, trans=full, acc=acc.astype(bool).tolist(), lookup=lookup,
                literal=lit, exact=exact, prefilter=pref, fixed=(ftw, fng, fcc, list(foff), list(fwid)))


PATTERNS = sorted({v["pattern"] for v in load_vectors()["vectors"]} | {
    r"(hello|world|foo)", r"[a-zA-Z0-9._%+-]+@[a-zA-Z0-9.-]+\.[a-z]{2,}", r"(\d{3})(\d{3})(\d{4})",
    r"\s*,\s*", r"(hello|help|helicopter)", r"(a|b)*", r"(cat|dog)+", r"(ab)?", r"(abc)+", r"(abc)*",
    r"a*", r"ab+", r"x?", r".+", r".?", r".", r"[0-9]+\.?[0-9]*", r"^[0-9]+\.?[0-9]*$", r"(?:00|33|44)[0-9]{2}",
    r"ab[0-9]{2}[a-z]{3}", r"\w+\s\w+", r"(\w+) (\w+)", r"[+]*\d+[-]*\d+", r"[abc]+", r"[^x]", r"a{2}b{3}",
    r".*", r"", r"^", r"$", r"^$", r"\t", r"a|b|c", r"(a|b)|(c|d)", r"[a-f0-9]{8}-[a-f0-9]{4}", r"\d{1,3}\.\d{1,3}",
    # the reference's LazyDFA route (NFAMatcher): subset construction in both front ends
    r"\s+", r"(\d+)", r"\d|\w", r"(\w\d)+", r"(a|b)x", r"\w*@", r"\s", r"3[02]|40", r"3[02]|40|[68]9",
    r"24|[346]|7[2-57-9]|9[2-9]", r"[^ab]+x", r"(a|ab)(c|bcd)", r"(ab)*c?", r"(?:hello) (\w+)", r"[\s\d]+x?", r"(x|y)*z{2,3}",
    r"h[ae]llo", r"(hello)", r"^[a-z]+$", r"hello.world", r"\d*abc", r"a.c", r".*x", r"x.*", r"(a|b){300}",
} | {v["pattern"] for v in load_vectors()["group_sub"]})


@pytest.mark.parametrize("pattern", PATTERNS, ids=lambda p: p or "<empty>")
def test_product_front_end_matches_oracle_tables(pattern):
    try:
        oc = oracle.frontend.compile_regex(pattern.encode())
    except oracle.Unsupported:
        with pytest.raises(mr.UnsupportedPattern):
            mr.CompiledRegex(pattern)
        return
    except oracle.RegexError:
        with pytest.raises(mr.RegexError):
            mr.CompiledRegex(pattern)
        return
    rx = mr.CompiledRegex(pattern)
    b = decode_blob(rx.blob())
    if oc.is_wildcard_match_any:
        assert b["flags"] & (1 << 5)
        return
    d = oc.dfa
    assert b["trans"].tolist() == d.trans
    assert b["acc"] == [bool(x) for x in d.accepting]
    assert b["start"] == d.start_state
    assert bool(b["flags"] & 1) == d.has_start_anchor and bool(b["flags"] & 2) == d.has_end_anchor
    assert bool(b["flags"] & 4) == d.is_pure_literal
    assert bool(b["flags"] & 8) == (d.matcher is not None)
    assert bool(b["flags"] & 16) == d.simd_scan_eligible
    assert bool(b["flags"] & 64) == oc.is_exact_literal
    assert bool(b["flags"] & 128) == oc.has_anchors_info
    assert bool(b["flags"] & 256) == (oc.prefilter_literal is not None)
    assert b["literal"] == d.literal_pattern
    assert b["req"] == oc.required_byte
    if d.matcher is not None:
        assert b["lookup"].tolist() == d.matcher.lookup
    if oc.prefilter_literal is not None:
        assert b["prefilter"] == oc.prefilter_literal
    assert b["fixed"][0] == oc.fixed_sub_total_width and b["fixed"][1] == oc.fixed_sub_num_groups
    assert bool(b["fixed"][2]) == oc.fixed_sub_concat
    if oc.fixed_sub_total_width >= 0:
        assert b["fixed"][3] == oc.fixed_sub_offsets and b["fixed"][4] == oc.fixed_sub_widths


def test_plans_for_the_five_configs():
    names = mr.MODE_NAMES
    i = mr.CompiledRegex(r"\d+").info()
    assert names[i.findall_mode] == "CLASS_RUNS" and i.findall_direct
    i = mr.CompiledRegex(r"(hello|world|foo)").info()
    assert names[i.findall_mode] == "EVERY_POS" and i.findall_direct and (i.n_states, i.n_classes) == (12, 9)
    assert i.findall_dense and i.search_dense  # literal set -> dense Aho-Corasick table
    i = mr.CompiledRegex(r"[a-zA-Z0-9._%+-]+@[a-zA-Z0-9.-]+\.[a-z]{2,}").info()
    assert names[i.search_mode] == "FIRST_CLASS" and names[i.findall_mode] == "REQUIRED_BYTE" and i.findall_direct
    i = mr.CompiledRegex(r"(\d{3})(\d{3})(\d{4})").info()
    assert names[i.search_mode] == "UNIFORM_RUN" and i.fixed_sub_total_width == 10
    i = mr.CompiledRegex(r"\s*,\s*").info()
    assert names[i.findall_mode] == "REQUIRED_BYTE" and i.findall_direct and i.required_byte == ord(",")
    # overlapping candidates -> selection pass needed
    assert not mr.CompiledRegex("aa").info().findall_direct and mr.CompiledRegex("aa").info().findall_dense
    assert not mr.CompiledRegex("a|ab|abc").info().findall_dense  # not prefix-free: longest-accept matters
    assert not mr.CompiledRegex(r"[a-z]+[0-9]+").info().findall_dense
    assert mr.CompiledRegex("ab").info().findall_direct
    assert not mr.CompiledRegex(r"[a-z]+[0-9]+").info().findall_direct


def test_blob_roundtrip_and_dfa_pack():
    rx = mr.CompiledRegex(r"[a-z]+[0-9]+")
    blob = rx.blob()
    rx2 = mr.CompiledRegex(blob=blob)
    assert rx2.blob() == blob
    # the Mojo-side serializer entry: raw 256 x int32 tables -> identical blob
    oc = oracle.frontend.compile_regex(rb"[a-z]+[0-9]+")
    d = oc.dfa
    desc = mr.DfaDesc()
    desc.n_states = len(d.trans)
    desc.start_state = d.start_state
    desc.scan_eligible = int(d.simd_scan_eligible)
    desc.required_byte = oc.required_byte
    desc.fixed_sub_total_width = oc.fixed_sub_total_width
    trans = np.asarray(d.trans, dtype=np.int32)
    acc = np.asarray(d.accepting, dtype=np.uint8)
    lookup = np.asarray(d.matcher.lookup, dtype=np.uint8)
    n = ctypes.c_size_t(0)
    assert mr.lib.mrx_dfa_pack(ctypes.byref(desc), trans.ctypes.data, acc.ctypes.data, lookup.ctypes.data,
                               None, 0, None, 0, None, 0, None, ctypes.byref(n)) == 0
    buf = ctypes.create_string_buffer(n.value)
    assert mr.lib.mrx_dfa_pack(ctypes.byref(desc), trans.ctypes.data, acc.ctypes.data, lookup.ctypes.data,
                               None, 0, None, 0, None, 0, ctypes.cast(buf, ctypes.c_void_p), ctypes.byref(n)) == 0
    assert buf.raw[:n.value] == blob
    # corrupt blobs are rejected
    with pytest.raises(ValueError):
        mr.CompiledRegex(blob=blob[:-1])
    bad = bytearray(blob)
    bad[0] ^= 0xFF
    with pytest.raises(ValueError):
        mr.CompiledRegex(blob=bytes(bad))


@pytest.mark.parametrize("pattern", [rb"(\d{3})(\d{3})(\d{4})", rb"(\d{3})-(\d{4})", rb"\s*,\s*", rb"(hello|world|foo)",
                                     rb"[a-zA-Z0-9._%+-]+@[a-zA-Z0-9.-]+\.[a-z]{2,}", rb"^abc", rb"hello"])
def test_serializer_entry_with_all_metadata_equals_the_compiled_blob(pattern):
    """mrx_dfa_pack fed with everything mojo/cuda_backend.mojo passes (raw rows, flags, required byte, literals,
    CompiledRegex._fixed_sub_*) gives byte for byte the blob mrx_compile builds -- in particular the fixed-width
    sub metadata of BASELINE config 4 survives the Mojo route."""
    from _pack import pack_through_the_serializer_entry

    blob = pack_through_the_serializer_entry(pattern)
    rx = mr.CompiledRegex(pattern)
    assert blob == rx.blob()
    a, b = mr.CompiledRegex(blob=blob).info(), rx.info()
    for k, _ in mr.Info._fields_:
        assert getattr(a, k) == getattr(b, k), k


def test_blob_with_bad_fixed_width_metadata_is_rejected():
    import struct

    blob = bytearray(mr.CompiledRegex(rb"(\d{3})(\d{3})(\d{4})").blob())
    off = 16 + 16 + 12          # magic..flags, n_states..required_byte, fixed_total_width / num_groups / concat
    for value in (-5, 1 << 20):
        bad = bytearray(blob)
        struct.pack_into("<i", bad, off + 4 * 2, value)      # fixed_offsets[2]
        with pytest.raises(ValueError):
            mr.CompiledRegex(blob=bytes(bad))
    bad = bytearray(blob)
    struct.pack_into("<i", bad, off + 40 + 4 * 3, 9)          # fixed_widths[3]: offset 6 + 9 > 10
    with pytest.raises(ValueError):
        mr.CompiledRegex(blob=bytes(bad))


def test_parse_errors_and_nfa_patterns_are_reported():
    for p in ["(abc", "abc)", "[abc", "a]", "a{x}"]:
        with pytest.raises(mr.RegexError):
            mr.CompiledRegex(p)
        with pytest.raises(oracle.RegexError):
            oracle.compile(p)
    for v in load_vectors()["nfa_routed"]:
        with pytest.raises(mr.UnsupportedPattern):
            mr.CompiledRegex(v["pattern"])


def test_matching_without_a_device_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(mr.RegexCudaError):
        mr.findall(r"\d+", b"abc123")
    # the entry points that take no text object fail the same way (no CPU fallback anywhere)
    import ctypes

    rx = mr.CompiledRegex(r"\d+")
    sp = mr._Spans()
    assert mr.lib.mrx_findall_async(rx._h, None, ctypes.byref(sp)) == -5       # MRX_E_NO_DEVICE
    assert mr.lib.mrx_findall_host(rx._h, None, 0, ctypes.byref(sp)) == -5
    assert mr.lib.mrx_comm_allgather_device(None, None) == -5
    assert mr.lib.mrx_init_multi(2) < 0 and mr.lib.mrx_device_count() == 0


def _random_patterns(n: int, seed: int):
    """Small grammar around the shapes compile_dfa_pattern knows (dfa.mojo:2385-2496): classes, literals,
    quantifiers, groups, alternations, anchors -- plus noise that should be rejected the same way by both
    front ends."""
    import random

    rng = random.Random(seed)
    atoms = [r"\d", r"\w", r"\s", "a", "b", "ab", "abc", "x", "-", r"\.", ",", "@", ".", "[a-z]", "[0-9]", "[A-Z]", "[a-c]",
             "[^0-9]", "[a-zA-Z0-9._%+-]", "[;,]", "[89]", "hello", "foo"]
    quants = ["", "", "", "+", "*", "?", "{2}", "{3}", "{2,}", "{1,3}", "{0,2}"]

    def piece():
        a = rng.choice(atoms)
        q = rng.choice(quants)
        if len(a) > 1 and a[0] not in "\\[" and q and rng.random() < 0.7:
            a = "(" + a + ")"
        return a + q

    out = []
    for _ in range(n):
        k = rng.choice([1, 1, 2, 2, 3, 4, 5])
        kind = rng.random()
        if kind < 0.15:
            p = "|".join(rng.choice(["a", "b", "ab", "abc", "hello", "world", "foo", "na", "nb", "cat", "dog"])
                         for _ in range(rng.randint(2, 4)))
            if rng.random() < 0.6:
                p = "(" + p + ")"
            if rng.random() < 0.3:
                p += rng.choice(["+", "*", "?"])
        elif kind < 0.3:
            p = "".join("(" + piece() + ")" for _ in range(k))
        else:
            p = "".join(piece() for _ in range(k))
        if rng.random() < 0.1:
            p = "^" + p
        if rng.random() < 0.1:
            p = p + "$"
        out.append(p)
    return sorted(set(out))


@pytest.mark.parametrize("pattern", _random_patterns(420, 20261018), ids=lambda p: p or "<empty>")
def test_random_patterns_front_ends_agree(pattern):
    """Differential test of the two independent restatements of the reference's front end (C++ in the
    product, Python in the oracle): same verdict (tables / unsupported / parse error) on random patterns."""
    test_product_front_end_matches_oracle_tables(pattern)


def test_search_skip_byte_is_read_off_the_automaton():
    """RevTable::req_byte (program.cpp: required_byte): the rarest byte every accepted string contains, or -1
    when there is none / only common ones.  Used by the batch kernel to skip documents."""
    cases = {
        r"[a-zA-Z0-9._%+-]+@[a-zA-Z0-9.-]+\.[a-zA-Z]{2,}": ord("@"),
        r"[A-Z]{2}-[0-9]{3}": ord("-"),
        r"[89]00[0-9]+": ord("0"),
        r"[a-z]+[0-9]+": -1,
        r"(hello|world|foo)": -1,   # 'o' is required but lower-case letters are too common to filter on
        r"a+b": -1,
    }
    for pattern, want in cases.items():
        assert mr.compile_regex(pattern).info().search_skip_byte == want, pattern


def test_skip_byte_is_required_according_to_the_oracle():
    """Whatever byte the planner derives as `search_skip_byte`, the restated reference must find nothing in
    documents that lack it -- and the filter must not be claimed for a byte some match can do without.
    Random DFA-routed patterns; documents drawn from the bytes the pattern mentions plus filler."""
    import random

    rng = random.Random(77)
    checked = 0
    for pattern in _random_patterns(420, 20261018):
        try:
            info = mr.compile_regex(pattern).info()
            orx = oracle.compile(pattern)
        except (mr.RegexError, mr.UnsupportedPattern, mr.UndefinedBehaviour, oracle.Unsupported, oracle.RegexError):
            continue
        c = info.search_skip_byte
        if c < 0:
            continue
        assert info.search_rev_states > 0
        alphabet = bytes(sorted(set(b"abcxhelowrdfn0189AZ.,;@-_%+ \t\n") - {c}))
        for _ in range(60):
            doc = bytes(rng.choice(alphabet) for _ in range(rng.randint(0, 48)))
            assert orx.search(doc) is None, (pattern, chr(c), doc)
        # and with the byte available the pattern is satisfiable at all (the filter is not vacuous)
        hit = False
        full = alphabet + bytes([c]) * 6
        for _ in range(4000):
            doc = bytes(rng.choice(full) for _ in range(rng.randint(1, 24)))
            if orx.search(doc) is not None:
                hit = True
                break
        checked += 1 if hit else 0
    assert checked >= 6


def test_leading_run_family_is_recognised_only_where_the_argument_holds():
    """compute_leading_run (api.cu): findall attempts only the first byte of a run of X when the pattern begins with
    X+ and never reads X again.  `a+1a?` begins with a+ but reads an `a` after the 1: "a1aa1" -> (0,3), (3,5), and
    the second match starts on an `a` whose predecessor is an `a` -- pruning it would lose the match."""
    yes = [r"[a-z]+[0-9]+", r"[ab]+c", r"\d+x", r"[a-z]+\d*"]
    # (`a+b`: the reference's builder makes 0 -a-> 2 -b-> 3 -a-> 1, no loop and an `a` read after the b;
    #  `[a-z]+@[0-9]+` goes through the required-byte plan, not through per-candidate attempts)
    no = [r"a+1a?", r"a+b", r"[a-z]+@[0-9]+", r"[a-z]+[a-z0-9]+x", r"ab+", r"[a-z]+", r"(\d{3})(\d{3})(\d{4})",
          r"hello", r"a*b"]
    for pat in yes:
        assert mr.CompiledRegex(pat).info().findall_leading_run == 1, pat
    for pat in no:
        assert mr.CompiledRegex(pat).info().findall_leading_run == 0, pat
