"""ORACLE (test infrastructure, NOT product code).

`oracle.compile(pattern)` -> OracleRegex with the reference's re-style operations
(search / match_first / findall / sub / split), computed by the CPU restatement:
frontend.py (pattern -> tables) + scan.c (text loops, via ctypes).

Importers allowed: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline and
--impl reference legs).  The product never imports this package.
"""

from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Optional, Tuple

import numpy as np

from . import frontend
from .frontend import RegexError, Unsupported  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_scan.so")


def build(force: bool = False) -> str:
    """Compile oracle/scan.c -> oracle/liboracle_scan.so (gcc -O3)."""
    src = os.path.join(_HERE, "scan.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O3", "-march=x86-64-v2", "-fPIC", "-shared", "-o", _SO, src])
    return _SO


class _CRegex(ctypes.Structure):
    _fields_ = [
        ("trans", ctypes.c_void_p), ("accepting", ctypes.c_void_p),
        ("n_states", ctypes.c_int32), ("start_state", ctypes.c_int32),
        ("has_start_anchor", ctypes.c_int32), ("has_end_anchor", ctypes.c_int32),
        ("is_pure_literal", ctypes.c_int32), ("has_matcher", ctypes.c_int32),
        ("scan_eligible", ctypes.c_int32),
        ("literal", ctypes.c_void_p), ("literal_len", ctypes.c_int64),
        ("lookup", ctypes.c_uint8 * 256), ("simd_cand", ctypes.c_uint8 * 256),
        ("simd_width", ctypes.c_int32),
        ("is_wildcard_any", ctypes.c_int32), ("is_exact_literal", ctypes.c_int32),
        ("has_anchors_info", ctypes.c_int32), ("required_byte", ctypes.c_int32),
        ("exact_literal", ctypes.c_void_p), ("exact_len", ctypes.c_int64),
        ("prefilter", ctypes.c_void_p), ("prefilter_len", ctypes.c_int64),
        ("fixed_total_width", ctypes.c_int32), ("fixed_num_groups", ctypes.c_int32),
        ("fixed_concat", ctypes.c_int32),
        ("fixed_offsets", ctypes.c_int32 * 10), ("fixed_widths", ctypes.c_int32 * 10),
        ("nib_lookup", ctypes.c_uint8 * 32), ("nib_simd", ctypes.c_uint8 * 32), ("use_avx2", ctypes.c_int32),
    ]


def _nibble_rows(table) -> list:
    """scan.c member32: per low nibble the membership bits of the high nibbles 0-7 ([0..15]) and 8-15 ([16..31])."""
    rows = [0] * 32
    for b in range(256):
        if table[b]:
            lo, hi = b & 15, b >> 4
            rows[lo + (16 if hi >= 8 else 0)] |= 1 << (hi & 7)
    return rows


class _Seg(ctypes.Structure):
    _fields_ = [("group_ref", ctypes.c_int32), ("start", ctypes.c_int32), ("length", ctypes.c_int32)]


_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        P = ctypes.POINTER
        lib.orx_free.argtypes = [ctypes.c_void_p]
        lib.orx_match_first.argtypes = [P(_CRegex), ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                        P(ctypes.c_int64)]
        lib.orx_match_first.restype = ctypes.c_int
        lib.orx_match_next.argtypes = lib.orx_match_first.argtypes
        lib.orx_match_next.restype = ctypes.c_int
        lib.orx_match_all.argtypes = [P(_CRegex), ctypes.c_void_p, ctypes.c_int64,
                                      P(ctypes.c_void_p)]
        lib.orx_match_all.restype = ctypes.c_int64
        lib.orx_sub.argtypes = [P(_CRegex), ctypes.c_char_p, ctypes.c_int64, ctypes.c_int, P(_Seg),
                                ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                P(ctypes.c_void_p), P(ctypes.c_int64)]
        lib.orx_sub.restype = ctypes.c_int
        lib.orx_split.argtypes = [P(_CRegex), ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                  P(ctypes.c_void_p), P(ctypes.c_void_p)]
        lib.orx_split.restype = ctypes.c_int64
        lib.orx_search_batch.argtypes = [P(_CRegex), ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                         ctypes.c_void_p]
        lib.orx_search_batch.restype = None
        _lib = lib
    return _lib


def _as_u8(text) -> np.ndarray:
    if isinstance(text, np.ndarray):
        assert text.dtype == np.uint8
        return np.ascontiguousarray(text)
    if isinstance(text, str):
        text = text.encode("utf-8")
    return np.frombuffer(bytes(text), dtype=np.uint8)


class OracleRegex:
    """CompiledRegex restatement (matcher.mojo:929-1163)."""

    def __init__(self, pattern, simd_width: int = 32):
        if isinstance(pattern, str):
            pattern = pattern.encode("utf-8")
        self.pattern = bytes(pattern)
        self.compiled = frontend.compile_regex(self.pattern)
        c = self.compiled
        cr = _CRegex()
        self._keep = []
        if c.dfa is not None:
            d = c.dfa
            trans = np.asarray(d.trans, dtype=np.int32).reshape(-1, 256)
            acc = np.asarray(d.accepting, dtype=np.uint8)
            lit = np.frombuffer(d.literal_pattern, dtype=np.uint8) if d.literal_pattern else np.zeros(1, np.uint8)
            self._keep += [trans, acc, lit]
            cr.trans = trans.ctypes.data
            cr.accepting = acc.ctypes.data
            cr.n_states = trans.shape[0]
            cr.start_state = d.start_state
            cr.has_start_anchor = int(d.has_start_anchor)
            cr.has_end_anchor = int(d.has_end_anchor)
            cr.is_pure_literal = int(d.is_pure_literal)
            cr.has_matcher = int(d.matcher is not None)
            cr.scan_eligible = int(d.simd_scan_eligible)
            cr.literal = lit.ctypes.data
            cr.literal_len = len(d.literal_pattern)
            if d.matcher is not None:
                for i in range(256):
                    cr.lookup[i] = d.matcher.lookup[i]
                    cr.simd_cand[i] = d.matcher.simd_candidates[i]
                for i, (a, b) in enumerate(zip(_nibble_rows(d.matcher.lookup), _nibble_rows(d.matcher.simd_candidates))):
                    cr.nib_lookup[i] = a
                    cr.nib_simd[i] = b
        cr.use_avx2 = 0 if os.environ.get("ORACLE_NO_AVX2") else 1
        cr.simd_width = simd_width
        cr.is_wildcard_any = int(c.is_wildcard_match_any)
        cr.is_exact_literal = int(c.is_exact_literal)
        cr.has_anchors_info = int(c.has_anchors_info)
        cr.required_byte = c.required_byte
        ex = np.frombuffer(c.exact_literal, dtype=np.uint8) if c.exact_literal else np.zeros(1, np.uint8)
        self._keep.append(ex)
        cr.exact_literal = ex.ctypes.data
        cr.exact_len = len(c.exact_literal)
        if c.prefilter_literal:
            pf = np.frombuffer(c.prefilter_literal, dtype=np.uint8)
            self._keep.append(pf)
            cr.prefilter = pf.ctypes.data
            cr.prefilter_len = len(c.prefilter_literal)
        cr.fixed_total_width = c.fixed_sub_total_width
        cr.fixed_num_groups = c.fixed_sub_num_groups
        cr.fixed_concat = int(c.fixed_sub_concat)
        for i in range(10):
            cr.fixed_offsets[i] = c.fixed_sub_offsets[i]
            cr.fixed_widths[i] = c.fixed_sub_widths[i]
        self._c = cr
        self._lib = _load()

    # -- CompiledRegex.match_first / match_next / match_all (matcher.mojo:1048-1089)
    def match_first_at(self, text, start: int = 0) -> Optional[Tuple[int, int]]:
        t = _as_u8(text)
        out = (ctypes.c_int64 * 2)()
        ok = self._lib.orx_match_first(ctypes.byref(self._c), t.ctypes.data, t.size, start, out)
        return (out[0], out[1]) if ok else None

    def match_next(self, text, start: int = 0) -> Optional[Tuple[int, int]]:
        t = _as_u8(text)
        out = (ctypes.c_int64 * 2)()
        ok = self._lib.orx_match_next(ctypes.byref(self._c), t.ctypes.data, t.size, start, out)
        return (out[0], out[1]) if ok else None

    def match_all(self, text) -> np.ndarray:
        """-> int64 array of shape (n, 2)."""
        t = _as_u8(text)
        ptr = ctypes.c_void_p()
        n = self._lib.orx_match_all(ctypes.byref(self._c), t.ctypes.data, t.size, ctypes.byref(ptr))
        if n == 0:
            if ptr.value:
                self._lib.orx_free(ptr)
            return np.zeros((0, 2), dtype=np.int64)
        arr = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_int64)), shape=(n * 2,)).copy()
        self._lib.orx_free(ptr)
        return arr.reshape(n, 2)

    # -- module-level API semantics (matcher.mojo:1325-1415, 1857-1917)
    def search(self, text) -> Optional[Tuple[int, int]]:
        return self.match_next(text, 0)

    def search_batch(self, text, stride: int) -> np.ndarray:
        """search() per fixed-size document, looped in C (the reference's callers loop in compiled Mojo)."""
        t = _as_u8(text)
        n_docs = (t.size + stride - 1) // stride
        out = np.empty((n_docs, 2), dtype=np.int64)
        self._lib.orx_search_batch(ctypes.byref(self._c), t.ctypes.data, t.size, stride, n_docs, out.ctypes.data)
        return out

    def match_first(self, text) -> Optional[Tuple[int, int]]:
        r = self.match_first_at(text, 0)
        return r if (r is not None and r[0] == 0) else None  # matcher.mojo:1411-1415

    def findall(self, text) -> np.ndarray:
        return self.match_all(text)

    def sub(self, repl, text, count: int = 0) -> bytes:
        if isinstance(repl, str):
            repl = repl.encode("utf-8")
        t = _as_u8(text)
        use_groups = frontend.has_group_refs(repl)
        segs = frontend.parse_repl_template(repl) if use_groups else []
        arr = (_Seg * max(1, len(segs)))()
        for i, (g, s, l) in enumerate(segs):
            arr[i].group_ref, arr[i].start, arr[i].length = g, s, l
        out_v = ctypes.c_void_p()
        out_n = ctypes.c_int64()
        rc = self._lib.orx_sub(ctypes.byref(self._c), repl, len(repl), int(use_groups), arr, len(segs),
                               t.ctypes.data, t.size, count, ctypes.byref(out_v), ctypes.byref(out_n))
        if rc == -2:
            raise Unsupported("general capture-group sub uses the NFA engine (matcher.mojo:1783-1821)")
        if t.size == 0:
            return b""
        res = np.ctypeslib.as_array(ctypes.cast(out_v, ctypes.POINTER(ctypes.c_uint8)), shape=(out_n.value,)).tobytes()
        self._lib.orx_free(out_v)
        return res

    def split_raw(self, text, maxsplit: int = 0) -> Tuple[bytes, np.ndarray]:
        """-> (concatenated piece bytes, int64 offsets[P+1])."""
        t = _as_u8(text)
        ob = ctypes.c_void_p()
        oo = ctypes.c_void_p()
        n = self._lib.orx_split(ctypes.byref(self._c), t.ctypes.data, t.size, maxsplit,
                                ctypes.byref(ob), ctypes.byref(oo))
        if n == -3:
            raise Unsupported("overlapping separator spans: text[prev_end:start] with prev_end > start "
                              "is undefined in the reference (SURVEY appendix B-3)")
        offs = np.ctypeslib.as_array(ctypes.cast(oo, ctypes.POINTER(ctypes.c_int64)), shape=(n + 1,)).copy()
        data = (np.ctypeslib.as_array(ctypes.cast(ob, ctypes.POINTER(ctypes.c_uint8)), shape=(int(offs[-1]),)).tobytes()
                if offs[-1] > 0 else b"")
        if ob.value:
            self._lib.orx_free(ob)
        self._lib.orx_free(oo)
        return data, offs

    def split(self, text, maxsplit: int = 0) -> List[bytes]:
        data, offs = self.split_raw(text, maxsplit)
        return [data[offs[i]:offs[i + 1]] for i in range(len(offs) - 1)]


def compile(pattern, simd_width: int = 32) -> OracleRegex:  # noqa: A001
    return OracleRegex(pattern, simd_width)
