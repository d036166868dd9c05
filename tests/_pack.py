"""Shared by the CPU and GPU suites: the Mojo serializer's route into the library."""
import numpy as np

import mojo_regex_b200 as mr
import oracle


def pack_through_the_serializer_entry(pattern: bytes) -> bytes:
    """What mojo/cuda_backend.mojo does: raw DFAEngine rows + HybridMatcher / CompiledRegex flags -> blob."""
    import ctypes

    oc = oracle.frontend.compile_regex(pattern)
    d = oc.dfa
    desc = mr.DfaDesc()
    desc.n_states = len(d.trans)
    desc.start_state = d.start_state
    desc.has_start_anchor = int(d.has_start_anchor)
    desc.has_end_anchor = int(d.has_end_anchor)
    desc.is_pure_literal = int(d.is_pure_literal)
    desc.scan_eligible = int(d.simd_scan_eligible)
    desc.required_byte = oc.required_byte
    desc.is_wildcard_any = int(oc.is_wildcard_match_any)
    desc.is_exact_literal = int(oc.is_exact_literal)
    desc.has_anchors_info = int(oc.has_anchors_info)
    desc.fixed_sub_total_width = oc.fixed_sub_total_width
    desc.fixed_sub_num_groups = oc.fixed_sub_num_groups
    desc.fixed_sub_concat = int(oc.fixed_sub_concat)
    for gi in range(10):
        desc.fixed_sub_offsets[gi] = oc.fixed_sub_offsets[gi]
        desc.fixed_sub_widths[gi] = oc.fixed_sub_widths[gi]
    trans = np.asarray(d.trans, dtype=np.int32)
    acc = np.asarray(d.accepting, dtype=np.uint8)
    lookup = np.asarray(d.matcher.lookup, dtype=np.uint8) if d.matcher is not None else None
    lit = d.literal_pattern or b""
    args = (ctypes.byref(desc), trans.ctypes.data, acc.ctypes.data, lookup.ctypes.data if lookup is not None else None,
            lit, len(lit), oc.exact_literal or b"", len(oc.exact_literal or b""),
            oc.prefilter_literal or b"", len(oc.prefilter_literal or b""))
    size = ctypes.c_size_t(0)
    mr._check(mr.lib.mrx_dfa_pack(*args, None, ctypes.byref(size)))
    blob = ctypes.create_string_buffer(size.value)
    mr._check(mr.lib.mrx_dfa_pack(*args, ctypes.cast(blob, ctypes.c_void_p), ctypes.byref(size)))
    return blob.raw[:size.value]
