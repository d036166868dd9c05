"""mojo_regex_b200 -- B200-native matching hot path behind mojo-regex's `re`-style API.

Host-side mirror (Python over ctypes) of the reference's public interface for the
matching path -- same names, argument meaning and error behaviour as
/root/reference/src/regex/matcher.mojo:1325-1415, 1857-1917:

    compile_regex(pattern) -> CompiledRegex
    search(pattern, text) / match_first(pattern, text) / findall(pattern, text)
    sub(pattern, repl, text, count=0) / split(pattern, text, maxsplit=0)

All matching runs in libmojoregex_cuda.so (hand-written sm_100a kernels); there is no
CPU fallback.  The library is loaded on first use; without the built library that raises ImportError;
matching without a CUDA device raises RegexCudaError.
"""

from __future__ import annotations

import ctypes
import os
from typing import Dict, List, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmojoregex_cuda.so")

MRX_OK, MRX_E_INVALID, MRX_E_PARSE, MRX_E_UNSUPPORTED, MRX_E_CUDA, MRX_E_NO_DEVICE, MRX_E_UNDEFINED = 0, -1, -2, -3, -4, -5, -6
OUT_HOST, OUT_DEVICE = 0, 1

MODE_NAMES = {0: "WILDCARD_ANY", 1: "ANCHORED", 2: "EXACT_LITERAL", 3: "LITERAL", 4: "CLASS_RUNS",
              5: "FIRST_CLASS", 6: "EVERY_POS", 7: "REQUIRED_BYTE", 8: "UNIFORM_RUN", -1: "UNSUPPORTED"}


class RegexError(Exception):
    """Pattern rejected by the parser (the reference `raises`, parser.mojo:225-237,340,399)."""


class UnsupportedPattern(Exception):
    """Pattern the reference serves from its NFA engines; the GPU path refuses it loudly."""


class UndefinedBehaviour(Exception):
    """Input on which the reference's own behaviour is undefined."""


class RegexCudaError(RuntimeError):
    """CUDA failure or missing device.  There is no CPU fallback."""


class _Spans(ctypes.Structure):
    _fields_ = [("start_end", ctypes.POINTER(ctypes.c_int64)), ("n", ctypes.c_size_t),
                ("location", ctypes.c_int), ("_owner", ctypes.c_void_p)]


class _Bytes(ctypes.Structure):
    _fields_ = [("bytes", ctypes.POINTER(ctypes.c_uint8)), ("n", ctypes.c_size_t),
                ("location", ctypes.c_int), ("_owner", ctypes.c_void_p)]


class _Pieces(ctypes.Structure):
    _fields_ = [("bytes", ctypes.POINTER(ctypes.c_uint8)), ("offsets", ctypes.POINTER(ctypes.c_int64)),
                ("n", ctypes.c_size_t), ("location", ctypes.c_int), ("_owner", ctypes.c_void_p)]


class Info(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int32) for k in (
        "n_states", "n_classes", "start_state", "has_start_anchor", "has_end_anchor", "is_pure_literal",
        "has_matcher", "scan_eligible", "required_byte", "is_wildcard_any", "is_exact_literal",
        "has_prefilter", "findall_mode", "search_mode", "findall_direct", "search_direct",
        "fixed_sub_total_width", "fixed_sub_num_groups", "fixed_sub_concat", "findall_dense", "search_dense",
        "search_rev_states", "search_skip_byte", "findall_sand", "split_fused", "search_via_prefilter",
        "sub_fused_width", "findall_leading_run", "search_dominated")]


class DfaDesc(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int32) for k in (
        "n_states", "start_state", "has_start_anchor", "has_end_anchor", "is_pure_literal", "scan_eligible",
        "required_byte", "is_wildcard_any", "is_exact_literal", "has_anchors_info",
        "fixed_sub_total_width", "fixed_sub_num_groups", "fixed_sub_concat")] + [
        ("fixed_sub_offsets", ctypes.c_int32 * 10), ("fixed_sub_widths", ctypes.c_int32 * 10)]


EXPORTS = [
    "mrx_abi_version", "mrx_init", "mrx_init_multi", "mrx_device_count", "mrx_shutdown", "mrx_last_error", "mrx_stream", "mrx_launch_count",
    "mrx_synchronize", "mrx_set_profiling", "mrx_last_kernel_ms", "mrx_kernel_ms_history", "mrx_compile", "mrx_dfa_pack", "mrx_regex_load", "mrx_regex_blob", "mrx_regex_info",
    "mrx_regex_free", "mrx_text_upload", "mrx_text_wrap_device", "mrx_text_set_window", "mrx_text_len",
    "mrx_text_free", "mrx_findall", "mrx_findall_async", "mrx_spans_count_device", "mrx_spans_wait", "mrx_findall_host", "mrx_findall_host_shard", "mrx_search", "mrx_match_first", "mrx_search_batch", "mrx_sub",
    "mrx_split", "mrx_sub_host", "mrx_split_host", "mrx_search_batch_host", "mrx_comm_unique_id", "mrx_comm_init", "mrx_comm_allgather", "mrx_comm_allgather_device", "mrx_comm_join",
    "mrx_comm_shutdown", "mrx_spans_free", "mrx_bytes_free", "mrx_pieces_free",
]


def _load() -> ctypes.CDLL:
    if not os.path.exists(_SO):
        raise ImportError(
            f"{_SO} is missing: build it with `python -m mojo_regex_b200.build` (nvcc, sm_100a). "
            "mojo_regex_b200 has no CPU fallback.")
    lib = ctypes.CDLL(_SO)
    P, c = ctypes.POINTER, ctypes
    vp, i64, sz = c.c_void_p, c.c_int64, c.c_size_t
    lib.mrx_abi_version.restype = c.c_int
    lib.mrx_init.argtypes = [c.c_int]
    lib.mrx_init_multi.argtypes = [c.c_int]
    lib.mrx_device_count.restype = c.c_int
    lib.mrx_last_error.restype = c.c_char_p
    lib.mrx_stream.restype = vp
    lib.mrx_launch_count.restype = c.c_uint64
    lib.mrx_set_profiling.argtypes = [c.c_int]
    lib.mrx_last_kernel_ms.restype = c.c_double
    lib.mrx_compile.argtypes = [c.c_char_p, sz, P(vp)]
    lib.mrx_dfa_pack.argtypes = [P(DfaDesc), vp, vp, vp, c.c_char_p, sz, c.c_char_p, sz, c.c_char_p, sz, vp, P(sz)]
    lib.mrx_regex_load.argtypes = [c.c_char_p, sz, P(vp)]
    lib.mrx_regex_blob.argtypes = [vp, vp, P(sz)]
    lib.mrx_regex_info.argtypes = [vp, P(Info)]
    lib.mrx_regex_free.argtypes = [vp]
    lib.mrx_regex_free.restype = None
    lib.mrx_text_upload.argtypes = [vp, sz, P(vp)]
    lib.mrx_text_wrap_device.argtypes = [vp, sz, P(vp)]
    lib.mrx_text_set_window.argtypes = [vp, sz, sz, i64]
    lib.mrx_text_len.argtypes = [vp]
    lib.mrx_text_len.restype = sz
    lib.mrx_text_free.argtypes = [vp]
    lib.mrx_text_free.restype = None
    lib.mrx_findall.argtypes = [vp, vp, c.c_int, P(_Spans)]
    lib.mrx_findall_async.argtypes = [vp, vp, P(_Spans)]
    lib.mrx_spans_count_device.argtypes = [P(_Spans)]
    lib.mrx_spans_count_device.restype = vp
    lib.mrx_spans_wait.argtypes = [P(_Spans)]
    lib.mrx_kernel_ms_history.argtypes = [P(c.c_double), c.c_int]
    lib.mrx_comm_allgather_device.argtypes = [vp, P(vp)]
    lib.mrx_findall_host.argtypes = [vp, vp, sz, P(_Spans)]
    lib.mrx_findall_host_shard.argtypes = [vp, vp, sz, sz, sz, i64, P(_Spans)]
    lib.mrx_search.argtypes = [vp, vp, i64, P(i64)]
    lib.mrx_match_first.argtypes = [vp, vp, i64, P(i64)]
    lib.mrx_search_batch.argtypes = [vp, vp, vp, sz, i64, c.c_int, vp]
    lib.mrx_sub.argtypes = [vp, c.c_char_p, sz, i64, vp, c.c_int, P(_Bytes)]
    lib.mrx_split.argtypes = [vp, vp, i64, c.c_int, P(_Pieces)]
    lib.mrx_sub_host.argtypes = [vp, c.c_char_p, sz, i64, vp, sz, P(_Bytes)]
    lib.mrx_split_host.argtypes = [vp, vp, sz, i64, P(_Pieces)]
    lib.mrx_search_batch_host.argtypes = [vp, vp, sz, vp, sz, i64, vp]
    lib.mrx_comm_unique_id.argtypes = [vp]
    lib.mrx_comm_init.argtypes = [vp, c.c_int, c.c_int]
    lib.mrx_comm_allgather.argtypes = [P(i64), c.c_int, P(vp)]
    lib.mrx_comm_join.argtypes = [vp]
    lib.mrx_spans_free.argtypes = [P(_Spans)]
    lib.mrx_spans_free.restype = None
    lib.mrx_bytes_free.argtypes = [P(_Bytes)]
    lib.mrx_bytes_free.restype = None
    lib.mrx_pieces_free.argtypes = [P(_Pieces)]
    lib.mrx_pieces_free.restype = None
    return lib


class _LazyLib:
    """libmojoregex_cuda.so is loaded on first use, not at import: helpers of this package that never touch
    the GPU (corpus generators, the shard arithmetic) can be imported without pulling the library in."""

    _lib = None

    def __getattr__(self, name):
        if _LazyLib._lib is None:
            _LazyLib._lib = _load()
        return getattr(_LazyLib._lib, name)


lib = _LazyLib()
_initialised_device: Optional[int] = None


def _check(rc: int) -> int:
    if rc >= 0:
        return rc
    msg = (lib.mrx_last_error() or b"").decode("utf-8", "replace")
    if rc == MRX_E_PARSE:
        raise RegexError(msg)
    if rc == MRX_E_UNSUPPORTED:
        raise UnsupportedPattern(msg)
    if rc == MRX_E_UNDEFINED:
        raise UndefinedBehaviour(msg)
    if rc == MRX_E_INVALID:
        raise ValueError(msg)
    raise RegexCudaError(f"[{rc}] {msg}")


def init(device: int = 0) -> None:
    """Select the CUDA device for this process (one process per GPU)."""
    global _initialised_device
    _check(lib.mrx_init(device))
    _initialised_device = device


def init_multi(n_gpus: int) -> None:
    """Multi-device mode: devices 0 .. n_gpus-1 in this one process; the host-buffer calls (match_all_host,
    search_batch_host, sub_host, split_host) then shard a large input over all of them inside one C-ABI call."""
    global _initialised_device
    _check(lib.mrx_init_multi(n_gpus))
    _initialised_device = 0


def device_count() -> int:
    return int(lib.mrx_device_count())


def shutdown() -> None:
    global _initialised_device
    _check(lib.mrx_shutdown())
    _initialised_device = None


def _ensure_init() -> None:
    if _initialised_device is None:
        init(int(os.environ.get("LOCAL_RANK", "0")))


def stream_handle() -> int:
    """cudaStream_t (as int) the library launches on."""
    _ensure_init()
    return int(lib.mrx_stream() or 0)


def launch_count() -> int:
    return int(lib.mrx_launch_count())


def set_profiling(enabled: bool) -> None:
    lib.mrx_set_profiling(int(enabled))


def last_kernel_ms() -> float:
    return float(lib.mrx_last_kernel_ms())


def kernel_ms_history(n: int):
    """Dominant-kernel times of the last n profiled calls, oldest first (synchronises once)."""
    buf = (ctypes.c_double * max(n, 1))()
    got = lib.mrx_kernel_ms_history(buf, n)
    return [float(buf[i]) for i in range(got)]


def synchronize() -> None:
    _check(lib.mrx_synchronize())


def comm_unique_id() -> bytes:
    """Rank 0: a fresh NCCL id (128 bytes) for comm_init on every rank."""
    buf = ctypes.create_string_buffer(128)
    _check(lib.mrx_comm_unique_id(ctypes.cast(buf, ctypes.c_void_p)))
    return buf.raw


def comm_init(uid: bytes, rank: int, world: int) -> None:
    """Join the library's own NCCL communicator (count exchange inside the library, no Python per step)."""
    _ensure_init()
    buf = ctypes.create_string_buffer(uid, 128)
    _check(lib.mrx_comm_init(ctypes.cast(buf, ctypes.c_void_p), rank, world))


def comm_allgather(values) -> int:
    """Queue one all-gather of a record of 1..4 int64; returns the device address of the (world, width) table."""
    arr = (ctypes.c_int64 * len(values))(*[int(v) for v in values])
    out = ctypes.c_void_p()
    _check(lib.mrx_comm_allgather(arr, len(values), ctypes.byref(out)))
    return int(out.value or 0)


def comm_allgather_device(value_dev: int) -> int:
    """Queue one all-gather of ONE int64 that lives in device memory and is produced on the library stream (the count
    of an asynchronous findall); returns the device address of the (world, 1) table."""
    out = ctypes.c_void_p()
    _check(lib.mrx_comm_allgather_device(ctypes.c_void_p(value_dev), ctypes.byref(out)))
    return int(out.value or 0)


def comm_join(stream_handle_: int = 0) -> None:
    _check(lib.mrx_comm_join(stream_handle_ or None))


def _as_bytes(x) -> bytes:
    if isinstance(x, str):
        return x.encode("utf-8")
    return bytes(x)


class Text:
    """A text resident in HBM.  Build from host bytes / numpy uint8, or wrap a CUDA tensor."""

    def __init__(self, data=None, *, device_ptr: int = 0, nbytes: int = 0, keepalive=None):
        _ensure_init()
        self._h = ctypes.c_void_p()
        self._keep = keepalive
        if data is not None:
            if isinstance(data, np.ndarray):
                assert data.dtype == np.uint8
                arr = np.ascontiguousarray(data)
                self._n = arr.size
                _check(lib.mrx_text_upload(arr.ctypes.data, arr.size, ctypes.byref(self._h)))
                synchronize()  # arr may be temporary
            else:
                b = _as_bytes(data)
                self._n = len(b)
                buf = ctypes.create_string_buffer(b, len(b)) if b else None
                _check(lib.mrx_text_upload(ctypes.cast(buf, ctypes.c_void_p) if buf else None, len(b),
                                           ctypes.byref(self._h)))
                synchronize()
        else:
            self._n = nbytes
            _check(lib.mrx_text_wrap_device(device_ptr, nbytes, ctypes.byref(self._h)))

    @classmethod
    def from_tensor(cls, t) -> "Text":
        """Wrap a contiguous CUDA uint8 torch tensor without copying.  The library stream is non-blocking and
        not ordered after torch's streams (see mrx_text_wrap_device): the producer stream of `t` is
        synchronised here so that the kernels never read an unfinished buffer."""
        import torch

        assert t.is_cuda and t.is_contiguous() and t.element_size() == 1
        torch.cuda.current_stream(t.device).synchronize()
        return cls(device_ptr=t.data_ptr(), nbytes=t.numel(), keepalive=t)

    @classmethod
    def from_pinned(cls, ptr: int, nbytes: int) -> "Text":
        """Upload `nbytes` from a (pinned) host pointer; asynchronous on the library stream."""
        self = cls.__new__(cls)
        _ensure_init()
        self._h = ctypes.c_void_p()
        self._keep = None
        self._n = nbytes
        _check(lib.mrx_text_upload(ptr, nbytes, ctypes.byref(self._h)))
        return self

    def set_window(self, core_begin: int, core_end: int, global_offset: int = 0) -> None:
        _check(lib.mrx_text_set_window(self._h, core_begin, core_end, global_offset))

    def __len__(self) -> int:
        return self._n

    def close(self) -> None:
        if self._h:
            lib.mrx_text_free(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Match:
    """matching.mojo:5-46 -- group_id 0, byte offsets, end-exclusive."""
    __slots__ = ("group_id", "start_idx", "end_idx", "_text")

    def __init__(self, start: int, end: int, text: Optional[bytes] = None):
        self.group_id = 0
        self.start_idx = int(start)
        self.end_idx = int(end)
        self._text = text

    def get_match_text(self) -> bytes:
        return self._text[self.start_idx:self.end_idx]

    def span(self):
        return (self.start_idx, self.end_idx)

    def __repr__(self):
        return f"Match(start_idx={self.start_idx}, end_idx={self.end_idx})"

    def __eq__(self, other):
        return isinstance(other, Match) and self.span() == other.span()


class MatchList:
    """matching.mojo:49-157 -- spans held as one (n, 2) int64 array (host numpy or CUDA tensor)."""

    def __init__(self, spans, text: Optional[bytes] = None, owner=None):
        self.spans = spans
        self._text = text
        self._owner = owner

    def __len__(self) -> int:
        return int(self.spans.shape[0])

    def __getitem__(self, i: int) -> Match:
        s, e = self.spans[i]
        return Match(int(s), int(e), self._text)

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def tolist(self):
        return [[int(s), int(e)] for s, e in self.spans]


def _text_of(text) -> "tuple[Text, Optional[bytes]]":
    if isinstance(text, Text):
        return text, None
    if isinstance(text, np.ndarray):
        return Text(text), None
    b = _as_bytes(text)
    return Text(b), b


class CompiledRegex:
    """matcher.mojo:929-1163."""

    def __init__(self, pattern=None, *, blob: Optional[bytes] = None):
        self._h = ctypes.c_void_p()
        if blob is not None:
            self.pattern = None
            _check(lib.mrx_regex_load(blob, len(blob), ctypes.byref(self._h)))
        else:
            self.pattern = _as_bytes(pattern)
            _check(lib.mrx_compile(self.pattern, len(self.pattern), ctypes.byref(self._h)))

    # -- introspection
    def info(self) -> Info:
        o = Info()
        _check(lib.mrx_regex_info(self._h, ctypes.byref(o)))
        return o

    def blob(self) -> bytes:
        n = ctypes.c_size_t(0)
        _check(lib.mrx_regex_blob(self._h, None, ctypes.byref(n)))
        buf = ctypes.create_string_buffer(n.value)
        _check(lib.mrx_regex_blob(self._h, ctypes.cast(buf, ctypes.c_void_p), ctypes.byref(n)))
        return buf.raw[:n.value]

    def get_stats(self) -> str:
        i = self.info()
        return (f"Pattern: {self.pattern!r}, Engine: DFA(cuda), Complexity: SIMPLE, "
                f"findall={MODE_NAMES[i.findall_mode]}{'/direct' if i.findall_direct else ''}, "
                f"search={MODE_NAMES[i.search_mode]}")

    # -- CompiledRegex.match_first / match_next / match_all
    def match_first(self, text, start: int = 0) -> Optional[Match]:
        t, raw = _text_of(text)
        out = (ctypes.c_int64 * 2)()
        rc = _check(lib.mrx_match_first(self._h, t._h, start, out))
        return Match(out[0], out[1], raw) if rc == 1 else None

    def match_next(self, text, start: int = 0) -> Optional[Match]:
        t, raw = _text_of(text)
        out = (ctypes.c_int64 * 2)()
        rc = _check(lib.mrx_search(self._h, t._h, start, out))
        return Match(out[0], out[1], raw) if rc == 1 else None

    def match_all(self, text) -> MatchList:
        if isinstance(text, np.ndarray) and text.dtype == np.uint8 and text.flags["C_CONTIGUOUS"]:
            _ensure_init()
            ml = self.match_all_host(text.ctypes.data, text.size)
            ml._keep = text
            return ml
        t, raw = _text_of(text)
        sp = _Spans()
        _check(lib.mrx_findall(self._h, t._h, OUT_HOST, ctypes.byref(sp)))
        try:
            if sp.n:
                arr = np.ctypeslib.as_array(sp.start_end, shape=(sp.n * 2,)).copy().reshape(-1, 2)
            else:
                arr = np.zeros((0, 2), dtype=np.int64)
        finally:
            lib.mrx_spans_free(ctypes.byref(sp))
        return MatchList(arr, raw)

    def match_all_host(self, ptr: int, nbytes: int, core_begin: int = 0, core_end: Optional[int] = None,
                       global_offset: int = 0) -> "MatchList":
        """findall over `nbytes` at host address `ptr` (page-locked for full speed): upload, scan and span
        download are pipelined inside the library.  The returned MatchList views the library's pinned
        result buffer (no copy); it is released when the MatchList is garbage collected."""
        _ensure_init()
        sp = _Spans()
        _check(lib.mrx_findall_host_shard(self._h, ptr, nbytes, core_begin, nbytes if core_end is None else core_end,
                                          global_offset, ctypes.byref(sp)))
        owner = _HostSpans(sp)
        if sp.n:
            arr = np.ctypeslib.as_array(sp.start_end, shape=(sp.n * 2,)).reshape(-1, 2)
        else:
            arr = np.zeros((0, 2), dtype=np.int64)
        return MatchList(arr, None, owner)

    def match_all_device(self, text: Text):
        """findall with the span array left in HBM: returns (torch int64 tensor view (n,2), keepalive)."""
        import torch

        sp = _Spans()
        _check(lib.mrx_findall(self._h, text._h, OUT_DEVICE, ctypes.byref(sp)))
        return _DeviceSpans(sp)

    def match_all_async(self, text: Text) -> "_DeviceSpans":
        """findall with the spans left in HBM and no host synchronisation: `.n` is -1 until `.wait()`;
        `.count_ptr` is the device address of the count (stream-ordered on the library stream)."""
        sp = _Spans()
        _check(lib.mrx_findall_async(self._h, text._h, ctypes.byref(sp)))
        return _DeviceSpans(sp)

    def search_batch(self, text, doc_offsets=None, doc_stride: int = 0, n_docs: Optional[int] = None,
                     out: Optional[np.ndarray] = None) -> np.ndarray:
        """One match_next(doc, 0) per document -> (n_docs, 2) int64, (-1,-1) = no match.  `out` may be a
        caller-provided (page-locked) int64 array of 2 * n_docs elements."""
        t, _ = _text_of(text)
        if doc_offsets is not None:
            offs = np.ascontiguousarray(doc_offsets, dtype=np.int64)
            n_docs = offs.size - 1
            optr = offs.ctypes.data
        else:
            assert doc_stride > 0
            if n_docs is None:
                n_docs = (len(t) + doc_stride - 1) // doc_stride
            optr = None
        if out is None:
            out = np.empty((n_docs, 2), dtype=np.int64)
        assert out.dtype == np.int64 and out.size >= 2 * n_docs and out.flags["C_CONTIGUOUS"]
        _check(lib.mrx_search_batch(self._h, t._h, optr, n_docs, doc_stride, OUT_HOST, out.ctypes.data))
        return out

    def search_batch_host(self, ptr: int, nbytes: int, out: np.ndarray, doc_offsets=None, doc_stride: int = 0,
                          n_docs: Optional[int] = None) -> np.ndarray:
        """Batch search over a host buffer (page-locked for full speed) into a caller-provided int64 array of
        2 * n_docs elements: document groups are uploaded, searched and downloaded as a pipeline."""
        _ensure_init()
        if doc_offsets is not None:
            offs = np.ascontiguousarray(doc_offsets, dtype=np.int64)
            n_docs = offs.size - 1
            optr = offs.ctypes.data
        else:
            assert doc_stride > 0
            if n_docs is None:
                n_docs = (nbytes + doc_stride - 1) // doc_stride
            optr = None
        assert out.dtype == np.int64 and out.size >= 2 * n_docs and out.flags["C_CONTIGUOUS"]
        _check(lib.mrx_search_batch_host(self._h, ptr, nbytes, optr, n_docs, doc_stride, out.ctypes.data))
        return out.reshape(-1)[: 2 * n_docs].reshape(n_docs, 2)

    def search_batch_device(self, text: "Text", out_ptr: int, n_docs: int, doc_stride: int = 0,
                            doc_offsets_ptr: int = 0, sync: bool = True) -> None:
        """Batch search with offsets / results already in HBM (out: n_docs x 2 int64).  The kernels run on
        the library stream; with sync=False the caller orders its own reads after them (stream_handle())."""
        _check(lib.mrx_search_batch(self._h, text._h, doc_offsets_ptr or None, n_docs, doc_stride, OUT_DEVICE, out_ptr))
        if sync:
            synchronize()

    def sub_device(self, repl, text: "Text", count: int = 0) -> "_DeviceBytes":
        r = _as_bytes(repl)
        ob = _Bytes()
        _check(lib.mrx_sub(self._h, r, len(r), count, text._h, OUT_DEVICE, ctypes.byref(ob)))
        return _DeviceBytes(ob)

    def split_device(self, text: "Text", maxsplit: int = 0) -> "_DevicePieces":
        pc = _Pieces()
        _check(lib.mrx_split(self._h, text._h, maxsplit, OUT_DEVICE, ctypes.byref(pc)))
        return _DevicePieces(pc)

    def sub(self, repl, text, count: int = 0) -> bytes:
        t, _ = _text_of(text)
        r = _as_bytes(repl)
        ob = _Bytes()
        _check(lib.mrx_sub(self._h, r, len(r), count, t._h, OUT_HOST, ctypes.byref(ob)))
        try:
            return np.ctypeslib.as_array(ob.bytes, shape=(ob.n,)).tobytes() if ob.n else b""
        finally:
            lib.mrx_bytes_free(ctypes.byref(ob))

    def sub_view(self, repl, text, count: int = 0) -> "HostBytes":
        """sub() whose result stays in the library's pinned buffer: .array is a zero-copy uint8 view."""
        t, _ = _text_of(text)
        r = _as_bytes(repl)
        ob = _Bytes()
        _check(lib.mrx_sub(self._h, r, len(r), count, t._h, OUT_HOST, ctypes.byref(ob)))
        return HostBytes(ob)

    def split_view(self, text, maxsplit: int = 0) -> "HostPieces":
        """split() whose pieces stay in the library's pinned buffers: .data (uint8) and .offsets (int64,
        n + 1 entries) are zero-copy views; piece i is data[offsets[i]:offsets[i+1]]."""
        t, _ = _text_of(text)
        pc = _Pieces()
        _check(lib.mrx_split(self._h, t._h, maxsplit, OUT_HOST, ctypes.byref(pc)))
        return HostPieces(pc)

    def sub_host(self, repl, ptr: int, nbytes: int, count: int = 0) -> "HostBytes":
        """sub() over `nbytes` at host address `ptr` (page-locked for full speed): chunked upload / rewrite /
        download pipeline inside the library; the result stays in the library's pinned buffer."""
        _ensure_init()
        r = _as_bytes(repl)
        ob = _Bytes()
        _check(lib.mrx_sub_host(self._h, r, len(r), count, ptr, nbytes, ctypes.byref(ob)))
        return HostBytes(ob)

    def split_host(self, ptr: int, nbytes: int, maxsplit: int = 0) -> "HostPieces":
        """split() over a host buffer, pipelined like sub_host; pieces stay in the library's pinned buffers."""
        _ensure_init()
        pc = _Pieces()
        _check(lib.mrx_split_host(self._h, ptr, nbytes, maxsplit, ctypes.byref(pc)))
        return HostPieces(pc)

    def split_raw(self, text, maxsplit: int = 0):
        t, _ = _text_of(text)
        pc = _Pieces()
        _check(lib.mrx_split(self._h, t._h, maxsplit, OUT_HOST, ctypes.byref(pc)))
        try:
            offs = np.ctypeslib.as_array(pc.offsets, shape=(pc.n + 1,)).copy()
            data = np.ctypeslib.as_array(pc.bytes, shape=(int(offs[-1]),)).tobytes() if offs[-1] else b""
        finally:
            lib.mrx_pieces_free(ctypes.byref(pc))
        return data, offs

    def split(self, text, maxsplit: int = 0) -> List[bytes]:
        data, offs = self.split_raw(text, maxsplit)
        return [data[offs[i]:offs[i + 1]] for i in range(len(offs) - 1)]

    def close(self):
        if self._h:
            lib.mrx_regex_free(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HostBytes:
    """sub() output in the library's pinned host memory (released on close / garbage collection)."""

    def __init__(self, ob: _Bytes):
        self._ob = ob
        self.n = int(ob.n)
        self.array = np.ctypeslib.as_array(ob.bytes, shape=(self.n,)) if self.n else np.zeros(0, dtype=np.uint8)

    def tobytes(self) -> bytes:
        return self.array.tobytes()

    def close(self):
        if self._ob is not None:
            self.array = None
            lib.mrx_bytes_free(ctypes.byref(self._ob))
            self._ob = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HostPieces:
    """split() output in the library's pinned host memory."""

    def __init__(self, pc: _Pieces):
        self._pc = pc
        self.n = int(pc.n)
        self.offsets = np.ctypeslib.as_array(pc.offsets, shape=(self.n + 1,))
        total = int(self.offsets[-1])
        self.data = np.ctypeslib.as_array(pc.bytes, shape=(total,)) if total else np.zeros(0, dtype=np.uint8)

    def piece(self, i: int) -> bytes:
        return self.data[self.offsets[i]:self.offsets[i + 1]].tobytes()

    def close(self):
        if self._pc is not None:
            self.data = self.offsets = None
            lib.mrx_pieces_free(ctypes.byref(self._pc))
            self._pc = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _HostSpans:
    """Keeps a pinned host span array of the library alive for the numpy view built on it."""

    def __init__(self, sp: _Spans):
        self._sp = sp

    def close(self):
        if self._sp is not None:
            lib.mrx_spans_free(ctypes.byref(self._sp))
            self._sp = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _DeviceSpans:
    """Span array living in HBM (owned by the library until closed)."""

    def __init__(self, sp: _Spans):
        self._sp = sp
        self.n = -1 if sp.n == ctypes.c_size_t(-1).value else int(sp.n)
        self.ptr = ctypes.cast(sp.start_end, ctypes.c_void_p).value or 0

    @property
    def count_ptr(self) -> int:
        """Device address of the span count (uint64), valid in stream order on the library stream."""
        return int(lib.mrx_spans_count_device(ctypes.byref(self._sp)) or 0)

    def wait(self) -> bool:
        """Fetch the count of an asynchronous findall.  False: the span array was too small (repeat the call)."""
        rc = _check(lib.mrx_spans_wait(ctypes.byref(self._sp)))
        self.n = int(self._sp.n)
        return rc == 0

    @property
    def __cuda_array_interface__(self):
        return {"shape": (self.n, 2), "typestr": "<i8", "data": (self.ptr, False), "version": 3, "strides": None}

    def tensor(self):
        """Zero-copy torch view (n, 2) int64 of the spans in HBM (valid until close())."""
        import torch

        if self.n == 0:
            return torch.zeros((0, 2), dtype=torch.int64, device="cuda")
        return torch.as_tensor(self, device="cuda")

    def to_numpy(self) -> np.ndarray:
        return self.tensor().cpu().numpy()

    def close(self):
        if self._sp is not None:
            lib.mrx_spans_free(ctypes.byref(self._sp))
            self._sp = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _DeviceBytes:
    """sub() output left in HBM."""

    def __init__(self, ob: _Bytes):
        self._ob = ob
        self.n = int(ob.n)
        self.ptr = ctypes.cast(ob.bytes, ctypes.c_void_p).value or 0

    @property
    def __cuda_array_interface__(self):
        return {"shape": (self.n,), "typestr": "|u1", "data": (self.ptr, False), "version": 3, "strides": None}

    def tensor(self):
        import torch

        if self.n == 0:
            return torch.zeros((0,), dtype=torch.uint8, device="cuda")
        return torch.as_tensor(self, device="cuda")

    def close(self):
        if self._ob is not None:
            lib.mrx_bytes_free(ctypes.byref(self._ob))
            self._ob = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _DevicePieces:
    """split() output left in HBM: n pieces, bytes + (n+1) int64 offsets."""

    def __init__(self, pc: _Pieces):
        self._pc = pc
        self.n = int(pc.n)
        self.bytes_ptr = ctypes.cast(pc.bytes, ctypes.c_void_p).value or 0
        self.offsets_ptr = ctypes.cast(pc.offsets, ctypes.c_void_p).value or 0

    def offsets_tensor(self):
        import torch

        class _V:
            pass

        v = _V()
        v.__cuda_array_interface__ = {"shape": (self.n + 1,), "typestr": "<i8", "data": (self.offsets_ptr, False),
                                      "version": 3, "strides": None}
        return torch.as_tensor(v, device="cuda")

    def close(self):
        if self._pc is not None:
            lib.mrx_pieces_free(ctypes.byref(self._pc))
            self._pc = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- module-level API with the reference's pattern cache (matcher.mojo:1166-1321) ---------------
_cache: Dict[bytes, CompiledRegex] = {}


def compile_regex(pattern) -> CompiledRegex:
    key = _as_bytes(pattern)
    rx = _cache.get(key)
    if rx is None:
        rx = CompiledRegex(key)
        _cache[key] = rx
    return rx


def clear_regex_cache() -> None:
    _cache.clear()


def search(pattern, text) -> Optional[Match]:
    return compile_regex(pattern).match_next(text, 0)


def findall(pattern, text) -> MatchList:
    return compile_regex(pattern).match_all(text)


def match_first(pattern, text) -> Optional[Match]:
    r = compile_regex(pattern).match_first(text, 0)
    return r if (r is not None and r.start_idx == 0) else None  # matcher.mojo:1411-1415


def sub(pattern, repl, text, count: int = 0) -> bytes:
    return compile_regex(pattern).sub(repl, text, count)


def split(pattern, text, maxsplit: int = 0) -> List[bytes]:
    return compile_regex(pattern).split(text, maxsplit)
