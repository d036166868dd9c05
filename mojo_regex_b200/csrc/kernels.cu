// Hand-written sm_100a kernels of the matching hot path.
//
//  k_scan<FILT>      candidate filter (SWAR, 16-B shared-memory vectors) + per-candidate attempt
//                    (dense DFA walk from a shared-memory table / literal compare / required-byte
//                    back-scan), ordered single-pass compaction by decoupled look-back.  Serves
//                    findall (direct or potential list), search (global first match) and batch
//                    search (per-document first match).
//  k_runs<FILT>      maximal class runs (CLASS_RUNS) and closed-form chunking of runs for chain
//                    automata over one class (UNIFORM_RUN), ordered compaction as above.
//  k_tile_trailing / k_carry_scan   run-length carry across tiles for UNIFORM_RUN.
//  k_select_*        greedy non-overlap selection over a potential-match list.
//  k_seg_*           output assembly for sub / split.
//
// Text tiles are staged into shared memory with the TMA 1-D bulk copy (cp.async.bulk, mbarrier
// completion); each thread then owns 144 contiguous bytes (9 x LDS.128, bank-conflict free).
#include "kernels.cuh"

namespace mrx {

// ---------------------------------------------------------------------------------------------
// shared-memory layout helpers
struct TextView {
  const uint8_t *g;   // whole buffer in HBM
  const uint8_t *s;   // staged window in shared memory
  int64_t s_begin, s_end;
  __device__ __forceinline__ uint8_t at(int64_t q) const {
    return (q >= s_begin && q < s_end) ? s[q - s_begin] : g[q];
  }
};

__device__ __forceinline__ void load_prog(DevProg *dst, const DevProg *src) {
  // header + the four 256-byte tables + literal, then only the live part of the transition table
  const int fixed = offsetof(DevProg, trans);
  const int total = fixed + ((src->trans_len + 15) & ~15);
  const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
  uint4 *d4 = reinterpret_cast<uint4 *>(dst);
  for (int i = threadIdx.x; i < total / 16; i += blockDim.x) d4[i] = s4[i];
}

// Stage text[begin, begin+bytes) into shared memory.  TMA bulk copy when the global address is
// 16-byte aligned (always true for library-owned texts), plain loads otherwise.
__device__ __forceinline__ void stage_tile(uint8_t *sm, const uint8_t *g, int64_t begin, int bytes, uint64_t *bar) {
  const bool aligned = ((reinterpret_cast<uintptr_t>(g + begin) & 15) == 0);
  const int bulk = aligned ? (bytes & ~15) : 0;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_proxy_async();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (bulk > 0) {
      mbar_expect_tx(bar, (uint32_t)bulk);
      tma_bulk_g2s(sm, g + begin, (uint32_t)bulk, bar);
    }
  }
  for (int i = bulk + threadIdx.x; i < bytes; i += blockDim.x) sm[i] = g[begin + i];
  if (bulk > 0) mbar_wait(bar, 0);
  __syncthreads();
}

template <int FILT>
__device__ __forceinline__ uint32_t filter_word(uint32_t x, const Filter &f, const uint8_t *cand) {
  if (FILT == FILT_NEEDLES) {
    // inexact-but-never-missing zero-byte test; candidates are re-checked against cand[]
    uint32_t t0 = x ^ f.n0, r = (t0 - kL) & ~t0;
    if (f.n_needles > 1) { uint32_t t = x ^ f.n1; r |= (t - kL) & ~t; }
    if (f.n_needles > 2) { uint32_t t = x ^ f.n2; r |= (t - kL) & ~t; }
    if (f.n_needles > 3) { uint32_t t = x ^ f.n3; r |= (t - kL) & ~t; }
    return r & kH;
  } else if (FILT == FILT_RANGE) {
    return swar_range(x, f.n0, f.n1, f.hi_is_ff != 0);
  } else if (FILT == FILT_LUT) {
    uint32_t r = 0;
    r |= cand[x & 255] ? 0x80u : 0;
    r |= cand[(x >> 8) & 255] ? 0x8000u : 0;
    r |= cand[(x >> 16) & 255] ? 0x800000u : 0;
    r |= cand[x >> 24] ? 0x80000000u : 0;
    return r;
  } else {
    return kH;
  }
}

// ---------------------------------------------------------------------------------------------
// attempts: DFAEngine._try_match_at_position (dfa.mojo:1905-2026) on the packed table
struct Attempt {
  const DevProg *p;
  TextView tv;
  int64_t lo, n;  // the text the reference would see is [lo, n)

  // _try_match_simd (dfa.mojo:2132-2197); returns end or -1 (fall through to the DFA walk)
  __device__ __forceinline__ int64_t simd_shortcut(int64_t start) const {
    int64_t q = start;
    while (q < n && p->lookup[tv.at(q)]) ++q;
    if (q == start && !(p->flags & DF_START_ACCEPTING)) return -1;
    if ((p->flags & DF_END_ANCHOR) && q != n) return -1;
    return q;
  }
  // sixteen bytes at q can be read from HBM with one aligned load (q is past the staged window)
  __device__ __forceinline__ bool global16(int64_t q) const {
    const bool ok = q + 16 <= n && (q < tv.s_begin || q >= tv.s_end) && (reinterpret_cast<uintptr_t>(tv.g + q) & 15) == 0;
    // one thread streaming through HBM: pull the lines 4 KB ahead into L2 (the walk is latency-bound)
    if (ok && (q & 127) == 0 && q + 4096 + 128 <= n) asm volatile("prefetch.global.L2 [%0];" ::"l"(tv.g + q + 4096));
    return ok;
  }
  __device__ __forceinline__ int64_t dfa(int64_t start) const {
    if (p->flags & DF_SHORTCUT) {
      int64_t e = simd_shortcut(start);
      if (e >= 0) return e;
    }
    int64_t last = (p->flags & DF_START_ACCEPTING) ? start : -1;
    int64_t q = start;
    if (p->flags & DF_DENSE_FWD) {
      // one lookup per byte: row offset -> next row offset (0 = dead, bit 2 = accepting)
      const uint8_t *fwd = p->trans;
      uint32_t s = (uint32_t)p->fwd_start;
      if (start >= tv.s_begin && start < tv.s_end) {  // walk inside the staged tile with 32-bit indices
        const int lim = (int)((n < tv.s_end ? n : tv.s_end) - tv.s_begin);
        int r = (int)(start - tv.s_begin), last_r = -1;
        while (r < lim) {
          s = *reinterpret_cast<const uint16_t *>(fwd + s + 2u * tv.s[r]);
          if (s == 0) break;
          ++r;
          if (s & 4u) last_r = r;
        }
        if (last_r >= 0) last = tv.s_begin + last_r;
        q = s == 0 ? n : tv.s_begin + r;
      }
      while (q < n) {  // continuation past the staged window (rare) or start outside it
        if (global16(q)) {
          // a state that loops on its input (X+ inside a run longer than the tile): sixteen independent lookups
          // instead of a chain of sixteen
          const uint4 w = *reinterpret_cast<const uint4 *>(tv.g + q);
          const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
          bool all = true;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            all = all && *reinterpret_cast<const uint16_t *>(fwd + s + 2u * ((ws[i >> 2] >> (8 * (i & 3))) & 255u)) == s;
          if (all) {
            q += 16;
            if (s & 4u) last = q;
            continue;
          }
        }
        s = *reinterpret_cast<const uint16_t *>(fwd + s + 2u * tv.at(q));
        if (s == 0) break;
        ++q;
        if (s & 4u) last = q;
      }
      if (last >= 0 && (p->flags & DF_END_ANCHOR) && last != n) return -1;
      return last;
    }
    int state = p->start_state;
    const int nc = p->n_classes;
    // fast path: walk inside the staged tile with 32-bit indices
    if (start >= tv.s_begin && start < tv.s_end) {
      const int lim = (int)((n < tv.s_end ? n : tv.s_end) - tv.s_begin);
      int r = (int)(start - tv.s_begin), last_r = -1;
      bool dead = false;
      while (r < lim) {
        const int nx = p->trans[state * nc + p->byte_class[tv.s[r]]];
        if (nx == 0xFF) { dead = true; break; }
        state = nx;
        ++r;
        if (p->accepting[state]) last_r = r;
      }
      if (last_r >= 0) last = tv.s_begin + last_r;
      q = tv.s_begin + r;
      if (dead) q = n;  // skip the slow continuation
    }
    while (q < n) {  // continuation past the staged window (rare) or start outside it
      if (global16(q)) {  // (as above: skip sixteen bytes on which the state loops)
        const uint4 w = *reinterpret_cast<const uint4 *>(tv.g + q);
        const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
        bool all = true;
#pragma unroll
        for (int i = 0; i < 16; ++i)
          all = all && p->trans[state * nc + p->byte_class[(ws[i >> 2] >> (8 * (i & 3))) & 255u]] == state;
        if (all) {
          q += 16;
          if (p->accepting[state]) last = q;
          continue;
        }
      }
      const int nx = p->trans[state * nc + p->byte_class[tv.at(q)]];
      if (nx == 0xFF) break;
      state = nx;
      ++q;
      if (p->accepting[state]) last = q;
    }
    if (last >= 0 && (p->flags & DF_END_ANCHOR) && last != n) return -1;
    return last;
  }
  __device__ __forceinline__ bool literal_at(int64_t pos) const {
    const int m = p->literal_len;
    if (pos + m > n) return false;
    for (int i = 0; i < m; ++i)
      if (tv.at(pos + i) != p->literal[i]) return false;
    return true;
  }
};

struct Hit {
  int64_t key, s, e;
};

// One candidate position -> at most one potential match.
__device__ __forceinline__ bool try_candidate(const ScanParams &prm, const Attempt &at, int64_t pos, Hit &h) {
  switch (prm.mode) {
    case MRX_MODE_LITERAL:
    case MRX_MODE_EXACT_LITERAL:
      if (!at.literal_at(pos)) return false;
      h.key = h.s = pos;
      h.e = pos + at.p->literal_len;
      return true;
    case MRX_MODE_REQUIRED_BYTE: {  // matcher.mojo:876-897
      int64_t start = pos;
      if (pos > at.tv.s_begin && pos <= at.tv.s_end) {
        // back-scan inside the staged tile with 32-bit indices; leaves the tile only for runs that reach its start
        const int floor_r = (int)((at.lo > at.tv.s_begin ? at.lo : at.tv.s_begin) - at.tv.s_begin);
        int r = (int)(pos - at.tv.s_begin);
        while (r > floor_r && at.p->lookup[at.tv.s[r - 1]]) --r;
        start = at.tv.s_begin + r;
        if (r > floor_r || start <= at.lo) goto scanned;  // stopped on a byte outside the class, or at the text start
      }
      while (start > at.lo && at.p->lookup[at.tv.at(start - 1)]) --start;
    scanned:
      if ((at.p->flags & DF_START_ANCHOR) && start > at.lo) return false;  // match_first, dfa.mojo:1866-1867
      int64_t e = at.dfa(start);
      if (e < 0 || e <= pos) return false;
      h.key = pos; h.s = start; h.e = e;
      return true;
    }
    default: {
      int64_t e = at.dfa(pos);
      if (e < 0) return false;
      h.key = h.s = pos;
      h.e = e;
      return true;
    }
  }
}

__device__ __forceinline__ int64_t doc_of(const ScanParams &prm, int64_t pos) {
  if (prm.doc_offsets == nullptr) return pos / prm.doc_stride;
  int64_t lo = 0, hi = prm.n_docs;  // last d with doc_offsets[d] <= pos
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (prm.doc_offsets[mid] <= pos) lo = mid; else hi = mid;
  }
  return lo;
}

// Layout of one CTA tile (kTile bytes): warp w owns the contiguous region [w*kWarpBytes, +kWarpBytes);
// in step v its lane t reads the 16-byte vector at w*kWarpBytes + 512*v + 16*t (coalesced, conflict
// free).  Candidate positions of a step are compacted into a per-warp queue (position order = lane
// order), then processed 32 at a time with every lane busy; matches are appended, still in position
// order, to a per-warp list that is flushed to HBM with coalesced 16-byte stores once the tile's
// global rank is known (decoupled look-back).
constexpr int kWarps = kThreads / 32;
constexpr int kWarpBytes = kTile / kWarps;  // 4608
constexpr int kSteps = kWarpBytes / 512;    // 9
constexpr int kQCap = 512;                  // candidates of one step: at most 32 lanes x 16 bytes
constexpr int kListCap = 224;               // buffered matches per warp and tile before the recompute path

struct ScanShared {
  DevProg prog;
  alignas(16) uint8_t text[kTile + kHalo];
  alignas(8) uint64_t bar;
  int scan_ws[kThreads / 32 + 2];
  int tile;
  int overflow;
  unsigned long long base;
  uint16_t q_pos[kWarps][kQCap];      // candidate position relative to the tile start
  uint32_t list_kb[kWarps][kListCap];  // (key - tile_start) | (key - start) << 16
  uint32_t list_len[kWarps][kListCap];
};

template <int FILT>
__global__ void __launch_bounds__(kThreads, 3) k_scan(const ScanParams prm) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  ScanShared &sh = *reinterpret_cast<ScanShared *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
    sh.tile = (prm.out_kind == OUT_SPANS) ? (int)atomicAdd(prm.ticket, 1u) : (int)blockIdx.x;
    sh.overflow = 0;
    // early exit for search: a match left of this tile is already known (block-uniform decision)
    sh.base = (prm.out_kind == OUT_FIRST) ? ld_volatile_u64(prm.best) : ~0ull;
  }
  load_prog(&sh.prog, prm.prog);
  __syncthreads();
  const int tile = sh.tile;
  const int64_t tile_start = prm.tile_origin + (int64_t)tile * kTile;
  if (prm.out_kind == OUT_FIRST && sh.base < (unsigned long long)tile_start) return;
  int stage_bytes = kTile + kHalo;
  if (tile_start + stage_bytes > prm.n) stage_bytes = (int)(prm.n - tile_start > 0 ? prm.n - tile_start : 0);
  stage_tile(sh.text, prm.text, tile_start, stage_bytes, &sh.bar);

  Attempt at;
  at.p = &sh.prog;
  at.tv.g = prm.text;
  at.tv.s = sh.text;
  at.tv.s_begin = tile_start;
  at.tv.s_end = tile_start + stage_bytes;
  at.lo = 0;
  at.n = prm.n;

  const int64_t key_lo = prm.key_lo, key_hi = prm.key_hi;  // candidate positions in [key_lo, key_hi)
  const uint8_t *cand = sh.prog.cand;
  uint16_t *q = sh.q_pos[warp];
  const int warp_rel0 = warp * kWarpBytes;

  // block-uniform: required-byte findall whose attempts can run on the staged tile alone (see the attempt loop)
  const bool lean_rb = prm.out_kind == OUT_SPANS && prm.mode == MRX_MODE_REQUIRED_BYTE && (sh.prog.flags & DF_DENSE_FWD) &&
                       !(sh.prog.flags & (DF_SHORTCUT | DF_START_ANCHOR | DF_END_ANCHOR | DF_START_ACCEPTING));
  int list_n = 0;                    // warp-uniform: matches found by this warp so far
  unsigned long long out_idx = 0;    // pass 1: global index of this warp's next match

  // PASS 0: find matches, buffer them per warp;  PASS 1 (only after an overflow): recompute, emit directly
  for (int pass = 0; pass < 2; ++pass) {
    list_n = 0;
    for (int v = 0; v < kSteps; ++v) {
      const int rel0 = warp_rel0 + 512 * v + 16 * lane;
      const int64_t pos0 = tile_start + rel0;
      // ---- 1. candidate mask of this lane's 16 bytes
      uint32_t m = 0;
      if (pos0 < key_hi && pos0 + 16 > key_lo) {
        const uint4 w = *reinterpret_cast<const uint4 *>(sh.text + rel0);
        const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int wi = 0; wi < 4; ++wi) {
          uint32_t fl;
          if (FILT == FILT_NEEDLES && prm.dominated) {  // the pruning below needs exact flags
            fl = swar_eq(ws[wi], prm.filt.n0);
            if (prm.filt.n_needles > 1) fl |= swar_eq(ws[wi], prm.filt.n1);
            if (prm.filt.n_needles > 2) fl |= swar_eq(ws[wi], prm.filt.n2);
            if (prm.filt.n_needles > 3) fl |= swar_eq(ws[wi], prm.filt.n3);
          } else {
            fl = filter_word<FILT>(ws[wi], prm.filt, cand);
          }
          m |= flags_to_nibble(fl) << (4 * wi);
        }
        if (pos0 + 16 > key_hi) m &= (1u << (int)(key_hi - pos0)) - 1u;
        if (pos0 < key_lo) m &= ~((1u << (int)(key_lo - pos0)) - 1u);
      }
      // dominated candidates (search / batch search; findall for the leading-run family, api.cu
      // compute_leading_run): a candidate whose predecessor byte is a candidate too can never start a reported
      // match -- drop it here, before the queue.  Document starts and the search start itself are kept.
      if (prm.dominated && m && (prm.out_kind != OUT_BATCH || prm.doc_offsets == nullptr)) {
        const uint32_t prev = (pos0 > 0 && cand[at.tv.at(pos0 - 1)]) ? 1u : 0u;
        uint32_t keep = 0;
        if (prm.out_kind == OUT_BATCH) {
          const int64_t r = pos0 % prm.doc_stride;
          const int64_t first = r == 0 ? 0 : prm.doc_stride - r;
          for (int64_t b = first; b < 16; b += prm.doc_stride) keep |= 1u << (int)b;
        } else if (prm.search_start >= pos0 && prm.search_start < pos0 + 16) {
          keep = 1u << (int)(prm.search_start - pos0);
        }
        m &= ~((m << 1) | prev) | keep;
      }
      if (prm.key_hi_inclusive_n && prm.n >= pos0 && prm.n < pos0 + 16) m |= 1u << (int)(prm.n - pos0);
      // ---- 2. compaction: exclusive scan of popcounts, lanes append their positions in order
      const int c = __popc(m);
      int off = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, off, o);
        if (lane >= o) off += t;
      }
      const int total = __shfl_sync(0xffffffffu, off, 31);
      off -= c;
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        q[off++] = (uint16_t)(rel0 + b);
      }
      __syncwarp();
      // ---- 3. attempts, 32 candidates at a time
      for (int j0 = 0; j0 < total; j0 += 32) {
        const int j = j0 + lane;
        bool hit = false;
        Hit h;
        h.key = h.s = h.e = 0;
        int64_t doc_lo = 0;
        int64_t cur_doc = 0;
        if (j < total) {
          const int rel = q[j];
          const int64_t pos = tile_start + rel;
          bool ok = true;
          if (FILT != FILT_ALL && pos < prm.n && !cand[sh.text[rel]]) ok = false;  // SWAR filter is a superset
          if (ok && lean_rb && pos < prm.n) {
            // required-byte attempt entirely inside the staged window, 32-bit indices, shared memory only:
            // back-scan over the matcher class (matcher.mojo:883-888), then the dense forward walk from
            // there (dfa.mojo:1979-2026); a match must end past the needle (matcher.mojo:891)
            const uint8_t *tx = sh.text, *lk = sh.prog.lookup, *fwd = sh.prog.trans;
            int st = rel;
            while (st > 0 && lk[tx[st - 1]]) --st;
            if (st > 0 || tile_start == 0) {  // else the run may continue left of the tile: generic path
              uint32_t sx = (uint32_t)sh.prog.fwd_start;
              int x = st, last = -1;
              while (x < stage_bytes) {
                sx = *reinterpret_cast<const uint16_t *>(fwd + sx + 2u * tx[x]);
                if (sx == 0) break;
                ++x;
                if (sx & 4u) last = x;
              }
              if (sx == 0 || tile_start + stage_bytes >= prm.n) {  // ended inside the window, or at the text's end
                ok = false;                                          // (handled here)
                if (last > rel) { hit = true; h.key = pos; h.s = tile_start + st; h.e = tile_start + last; }
              }
            }
          }
          if (ok && prm.out_kind == OUT_BATCH) {
            cur_doc = doc_of(prm, pos);
            doc_lo = prm.doc_offsets ? prm.doc_offsets[cur_doc] : cur_doc * prm.doc_stride;
            at.lo = doc_lo;
            at.n = prm.doc_offsets ? prm.doc_offsets[cur_doc + 1] : min(doc_lo + prm.doc_stride, prm.n);
          }
          if (ok && prm.out_kind == OUT_BATCH && prm.doc_min) {  // matcher.mojo:784-796: the search starts at the literal
            const unsigned long long lit = prm.doc_min[cur_doc];
            if (lit == ~0ull || (unsigned long long)(pos - doc_lo) < (lit >> 32)) ok = false;
          }
          if (ok && prm.dominated && pos > at.lo && pos > prm.search_start && pos < prm.n && cand[at.tv.at(pos - 1)]) ok = false;
          if (ok) hit = try_candidate(prm, at, pos, h);
        }
        if (prm.out_kind == OUT_FIRST) {
          if (hit) atomicMin(prm.best, (unsigned long long)h.key);
          continue;
        }
        if (prm.out_kind == OUT_BATCH) {
          if (hit) atomicMin(prm.doc_best + cur_doc, ((unsigned long long)(h.s - doc_lo) << 32) | (unsigned long long)(h.e - doc_lo));
          continue;
        }
        const unsigned hm = __ballot_sync(0xffffffffu, hit);
        const int r = list_n + __popc(hm & ((1u << lane) - 1u));
        if (hit) {
          if (pass == 0) {
            const int64_t back = h.key - h.s, len = h.e - h.s;
            if (r < kListCap && back <= 0xffff && len <= 0xffffffffll) {
              sh.list_kb[warp][r] = (uint32_t)(h.key - tile_start) | ((uint32_t)back << 16);
              sh.list_len[warp][r] = (uint32_t)len;
            } else {
              sh.overflow = 1;
            }
          } else {
            const unsigned long long o = out_idx + r;
            if (o < (unsigned long long)prm.cap) {
              st_span(prm.spans + 2 * o, h.s + prm.rebase, h.e + prm.rebase);
              if (prm.keys) prm.keys[o] = h.key + prm.rebase;
            }
          }
        }
        list_n += __popc(hm);
      }
      __syncwarp();
    }
    if (prm.out_kind != OUT_SPANS || pass == 1) return;

    // ---- tile epilogue: warp counts -> block scan -> look-back -> coalesced flush
    int total;
    const int rank = block_exclusive_scan(lane == 0 ? list_n : 0, sh.scan_ws, &total);
    if (tid < 32) {
      const unsigned long long b = lookback_exclusive(prm.desc, tile, (unsigned long long)total);
      if (tid == 0) {
        sh.base = b;
        if (tile == prm.n_tiles - 1) *prm.total = b + (unsigned long long)total;
      }
    }
    __syncthreads();
    out_idx = sh.base + (unsigned long long)__shfl_sync(0xffffffffu, rank, 0);
    if (!sh.overflow) {
      for (int i = lane; i < list_n; i += 32) {
        const unsigned long long o = out_idx + i;
        if (o < (unsigned long long)prm.cap) {
          const uint32_t kb = sh.list_kb[warp][i];
          const int64_t key = tile_start + (kb & 0xffffu);
          const int64_t s = key - (kb >> 16);
          st_span(prm.spans + 2 * o, s + prm.rebase, s + sh.list_len[warp][i] + prm.rebase);
          if (prm.keys) prm.keys[o] = key + prm.rebase;
        }
      }
      return;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// dense literal-set scan: one table lookup per text byte (Aho-Corasick completion of the reference's
// trie, see program.cpp:build_dense).  Thread t owns the 144 contiguous bytes [144 t, 144 t + 144) of
// the tile (9 x LDS.128, conflict free), warms its state up on the max_len - 1 bytes before them and
// then walks them; "a literal ends here" states are the highest-numbered ones, so the per-byte test is
// one compare.  Matches are parked in a per-warp pool tagged (lane, sequence) and written to their
// final, position-ordered slot once lane prefixes and the tile's global rank are known.
constexpr int kPriv = 12;  // private match records per thread and tile

struct DenseShared {
  alignas(16) uint8_t text[kTile + kHalo];
  alignas(8) uint64_t bar;
  int scan_ws[kThreads / 32 + 2];
  int tile;
  int overflow;
  unsigned long long base;
  unsigned long long lb_ws[2 * (kThreads / 32) + 2];
  // (end - chunk start) | len << 8 for every match of a thread, in position order; two generations so
  // that a tile's matches can be flushed while the next tile is already being loaded
  uint16_t priv[2][kThreads * kPriv];
  DenseHdr hdr;
  alignas(16) uint16_t table[8];    // the whole table follows
};

// Walk this thread's 144 bytes of the staged tile as TWO independent chains -- A = bytes [0, 80), B = bytes
// [80, 144), each warmed up on the max_len - 1 bytes before it -- interleaved instruction by instruction,
// so that every warp has two dependent shared-memory lookups in flight instead of one (with 8 warps per
// scheduler a single chain leaves the issue slots 37 % idle, profiles/r01_k_dense_g_ncu.txt).
// SPANS: tokens (end - chunk start) | len << 8 go to `mine`: chain A's fill it from slot 0 upwards, chain
// B's from slot kPriv - 1 downwards (one shared budget of kPriv); returns the number of matches, cnt_a =
// those of chain A, ovf = the budget was exceeded (the caller then takes the recompute path).  FIRST: atomicMin on best.  The packed state bytes (low byte of the
// row offset = row mod 32 and the accept flag) are written back over the consumed text, so the literal that
// ended at a hit is one byte load + one LUT load away instead of a replay of max_len dependent lookups.
// All threads of the block must call this (it contains a barrier).
constexpr int kChainA = 80;            // bytes of chain A (5 vectors); chain B takes the other 4

template <bool FIRST>
__device__ __forceinline__ int dense_scan_chunk(const DenseParams &prm, DenseShared &sh, const TextView &tv,
                                                int64_t tile_start, uint16_t *mine, int &cnt_a, bool &ovf) {
  const int tid = threadIdx.x;
  const uint8_t *tbl = reinterpret_cast<const uint8_t *>(sh.table);
  const int chunk_rel = tid * kChunk;
  // start positions are filtered against the key window only in tiles that touch its edges
  const bool edge_tile = tile_start - 256 < prm.key_lo || tile_start + kTile > prm.key_hi;
  const int valid = (int)min((int64_t)kTile, prm.n - tile_start);  // text bytes in this tile
  // warm-up: the state after the max_len - 1 bytes preceding a chain is exact
  uint32_t sa = 0, sb = 0;
  {
    const int w = sh.hdr.max_len - 1;
    // history bytes of a chain that lie in this tile: for w <= 16 they sit in the one aligned 16-byte vector
    // before the chain (one conflict-free LDS.128 instead of w byte loads that collide 4-way across lanes)
    auto warm16 = [&](const uint8_t *vec_end, uint32_t s0) {
      const uint4 v = *reinterpret_cast<const uint4 *>(vec_end - 16);
      const uint32_t ws[4] = {v.x, v.y, v.z, v.w};
      uint32_t st = s0;
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i >= 16 - w) st = *reinterpret_cast<const uint16_t *>(tbl + st + 2u * ((ws[i >> 2] >> (8 * (i & 3))) & 0xffu));
      return st;
    };
    if (tid > 0 && w <= 16) {
      if (chunk_rel <= valid) sa = warm16(sh.text + chunk_rel, 0u);
    } else if (tid > 0 && w <= kChunk) {
      if (chunk_rel <= valid)
        for (int r = chunk_rel - w; r < chunk_rel; ++r) sa = *reinterpret_cast<const uint16_t *>(tbl + sa + 2u * sh.text[r]);
    } else {
      const int64_t chunk0 = tile_start + chunk_rel;
      int64_t q = chunk0 - w;
      if (q < 0) q = 0;
      const int64_t stop = chunk0 < prm.n ? chunk0 : prm.n;
      for (; q < stop; ++q) sa = *reinterpret_cast<const uint16_t *>(tbl + sa + 2u * tv.at(q));
    }
    if (w <= 16) {
      sb = warm16(sh.text + chunk_rel + kChainA, 0u);
    } else if (w <= kChainA) {
      for (int r = chunk_rel + kChainA - w; r < chunk_rel + kChainA; ++r)
        sb = *reinterpret_cast<const uint16_t *>(tbl + sb + 2u * sh.text[r]);
    } else {  // literals longer than chain A: B's history reaches back past this chunk
      const int64_t b0 = tile_start + chunk_rel + kChainA;
      int64_t q = b0 - w;
      if (q < 0) q = 0;
      const int64_t stop = b0 < prm.n ? b0 : prm.n;
      for (; q < stop; ++q) sb = *reinterpret_cast<const uint16_t *>(tbl + sb + 2u * tv.at(q));
    }
  }
  __syncthreads();  // every warm-up has read its predecessor's tail: the text may now be overwritten
  int ca = 0, cb = 0;
  const bool whole_chunk = chunk_rel + kChunk <= valid;
  constexpr uint32_t kF = 0x04040404u;  // bit 2 of every packed byte = accept flag
  // one step of a chain: PRMT (extract), IMAD (address), LDS.U16 (next row), PRMT (pack the low byte)
#define MRX_DENSE_STEP(S, W, ACC, K)                                                                      \
  {                                                                                                       \
    const uint32_t b_ = __byte_perm(W, 0, 0x4440 + K);                                                    \
    uint32_t addr_;                                                                                       \
    asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(addr_) : "r"(b_), "r"(S));                                     \
    S = *reinterpret_cast<const uint16_t *>(tbl + addr_);                                                 \
    ACC = __byte_perm(ACC, S, K == 0 ? 0x3214 : K == 1 ? 0x3240 : K == 2 ? 0x3410 : 0x4210);              \
  }
  // hits of 32 consecutive bytes starting at chunk offset `base`; chain A stores upwards, chain B downwards
  auto take_hits = [&](uint32_t f, int base, bool is_b, int &cnt) {
    if (!whole_chunk) {  // only in the tile that holds the end of the text
      const int live = valid - (chunk_rel + base);  // bytes of these 32 that exist in the text
      if (live < 32) f = live <= 0 ? 0u : (f & ((1u << live) - 1u));
    }
    while (f) {
      const int i = __ffs(f) - 1;
      f &= f - 1;
      const int off = base + i;  // byte of the chunk the literal ends on
      const uint32_t len = sh.hdr.len_lut[sh.text[chunk_rel + off] >> 3];
      if (edge_tile || FIRST) {
        const int64_t start = tile_start + chunk_rel + off + 1 - (int64_t)len;
        if (start < prm.key_lo || start >= prm.key_hi) continue;
        if (FIRST) {
          atomicMin(prm.best, ((unsigned long long)start << 8) | len);
          continue;
        }
      }
      if (ca + cb < kPriv) mine[is_b ? kPriv - 1 - cnt : cnt] = (uint16_t)((off + 1) | (len << 8));
      ++cnt;
    }
  };
  auto mask8 = [&](const uint32_t *pk) {  // 8 packed words (32 state bytes) -> 32-bit hit mask
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) m |= (((pk[j] & kF) * 0x04081020u) >> 28) << (4 * j);
    return m;
  };
#pragma unroll 1
  for (int bt = 0; bt < 2; ++bt) {  // two double batches: A vectors 2bt, 2bt+1 with B vectors 5+2bt, 6+2bt
    const int ra = chunk_rel + 32 * bt, rb = chunk_rel + kChainA + 32 * bt;
    uint32_t pa[8], pb[8];
#pragma unroll
    for (int vv = 0; vv < 2; ++vv) {
      uint4 *slot_a = reinterpret_cast<uint4 *>(sh.text + ra + 16 * vv);
      uint4 *slot_b = reinterpret_cast<uint4 *>(sh.text + rb + 16 * vv);
      const uint4 a4 = *slot_a, b4 = *slot_b;
      const uint32_t wa[4] = {a4.x, a4.y, a4.z, a4.w}, wb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int wi = 0; wi < 4; ++wi) {
        uint32_t acc_a = 0, acc_b = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          MRX_DENSE_STEP(sa, wa[wi], acc_a, k)
          MRX_DENSE_STEP(sb, wb[wi], acc_b, k)
        }
        pa[4 * vv + wi] = acc_a;
        pb[4 * vv + wi] = acc_b;
      }
      *slot_a = make_uint4(pa[4 * vv], pa[4 * vv + 1], pa[4 * vv + 2], pa[4 * vv + 3]);
      *slot_b = make_uint4(pb[4 * vv], pb[4 * vv + 1], pb[4 * vv + 2], pb[4 * vv + 3]);
    }
    // (no "any flag set?" test first: with a hit every ~56 bytes some lane of the warp always has one)
    take_hits(mask8(pa), 32 * bt, false, ca);
    take_hits(mask8(pb), kChainA + 32 * bt, true, cb);
  }
  {  // the last vector of chain A (bytes 64..79) on its own
    uint4 *slot = reinterpret_cast<uint4 *>(sh.text + chunk_rel + 64);
    const uint4 a4 = *slot;
    const uint32_t wa[4] = {a4.x, a4.y, a4.z, a4.w};
    uint32_t pa[4];
#pragma unroll
    for (int wi = 0; wi < 4; ++wi) {
      uint32_t acc = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) MRX_DENSE_STEP(sa, wa[wi], acc, k)
      pa[wi] = acc;
    }
    *slot = make_uint4(pa[0], pa[1], pa[2], pa[3]);
    if ((pa[0] | pa[1] | pa[2] | pa[3]) & kF) {
      uint32_t m = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) m |= (((pa[j] & kF) * 0x04081020u) >> 28) << (4 * j);
      take_hits(m, 64, false, ca);
    }
  }
#undef MRX_DENSE_STEP
  cnt_a = ca;
  ovf = ca + cb > kPriv;
  return ca + cb;
}

// search variant: one CTA per tile, no output list
__global__ void __launch_bounds__(kThreads, 4) k_dense_first(const DenseParams prm) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  DenseShared &sh = *reinterpret_cast<DenseShared *>(smem_raw);
  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int64_t tile_start = prm.tile_origin + (int64_t)tile * kTile;
  if (tid == 0) sh.base = ld_volatile_u64(prm.best);
  {
    const int total = (int)sizeof(DenseHdr) + prm.table_bytes;
    const uint4 *s4 = reinterpret_cast<const uint4 *>(prm.dense);
    uint4 *d4 = reinterpret_cast<uint4 *>(&sh.hdr);
    for (int i = tid; i < total / 16; i += kThreads) d4[i] = s4[i];
  }
  __syncthreads();
  if (sh.base != ~0ull && (sh.base >> 8) < (unsigned long long)tile_start) return;
  int stage_bytes = kTile + kHalo;
  if (tile_start + stage_bytes > prm.n) stage_bytes = (int)(prm.n - tile_start > 0 ? prm.n - tile_start : 0);
  stage_tile(sh.text, prm.text, tile_start, stage_bytes, &sh.bar);
  TextView tv{prm.text, sh.text, tile_start, tile_start + stage_bytes};
  int ca;
  bool ovf;
  dense_scan_chunk<true>(prm, sh, tv, tile_start, nullptr, ca, ovf);
}

// findall variant: persistent CTAs.  Per iteration: take a ticket, start the TMA of that tile, and while
// it is in flight resolve the look-back of the PREVIOUS tile and flush its matches; then scan the new
// tile and publish its aggregate.  Nobody ever waits for a predecessor that is still scanning.
__global__ void __launch_bounds__(kThreads, 4) k_dense(const DenseParams prm) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  DenseShared &sh = *reinterpret_cast<DenseShared *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31;
  const bool text_aligned = (reinterpret_cast<uintptr_t>(prm.text + prm.tile_origin) & 15) == 0;
  const uint8_t *tbl = reinterpret_cast<const uint8_t *>(sh.table);
  if (tid == 0) {
    mbar_init(&sh.bar, 1);
    fence_proxy_async();
  }
  {  // automaton -> shared memory, once per CTA
    const int total = (int)sizeof(DenseHdr) + prm.table_bytes;
    const uint4 *s4 = reinterpret_cast<const uint4 *>(prm.dense);
    uint4 *d4 = reinterpret_cast<uint4 *>(&sh.hdr);
    for (int i = tid; i < total / 16; i += kThreads) d4[i] = s4[i];
  }
  __syncthreads();

  // state of the tile whose flush is pending
  bool pending = false;
  int p_tile = 0, p_cnt = 0, p_cnt_a = 0, p_rank = 0, p_total = 0, p_gen = 0;
  int64_t p_start = 0;

  auto take_ticket_and_load = [&]() {  // thread 0 only
    const int t = prm.tile_base + (int)atomicAdd(prm.ticket, 1u);
    sh.tile = t;
    if (t < prm.n_tiles) {
      const int64_t ts = prm.tile_origin + (int64_t)t * kTile;
      int64_t sb = prm.n - ts;
      if (sb > kTile + kHalo) sb = kTile + kHalo;
      if (sb < 0) sb = 0;
      const uint32_t bulk = text_aligned ? ((uint32_t)sb & ~15u) : 0u;
      mbar_expect_tx(&sh.bar, bulk);
      if (bulk) tma_bulk_g2s(sh.text, prm.text + ts, bulk, &sh.bar);
    }
  };
  auto flush_pending = [&]() {  // all threads; may wait for predecessors' aggregates (one scan behind)
    const unsigned long long p_base = lookback_resolve_block(prm.desc, p_tile, (unsigned long long)p_total, sh.lb_ws);
    if (tid == 0 && p_tile == prm.n_tiles - 1) *prm.total = p_base + (unsigned long long)p_total;
    const unsigned long long o = p_base + (unsigned long long)p_rank;
    const uint16_t *rec = sh.priv[p_gen] + tid * kPriv;
    const int64_t cbase = p_start + (int64_t)tid * kChunk + prm.rebase;
    const int room = o < (unsigned long long)prm.cap ? (int)min((unsigned long long)p_cnt, (unsigned long long)prm.cap - o) : 0;
    int64_t *dst = prm.spans + 2 * o;
    for (int i = 0; i < room; ++i, dst += 2) {
      const uint32_t r = rec[i < p_cnt_a ? i : kPriv - 1 - (i - p_cnt_a)];  // chain A's tokens, then chain B's
      const int64_t end = cbase + (r & 0xffu);
      st_span(dst, end - (int64_t)(r >> 8), end);
    }
    pending = false;
  };

  if (tid == 0) take_ticket_and_load();
  __syncthreads();
  int tile = sh.tile;
  for (int it = 0; tile < prm.n_tiles; ++it) {
    const int64_t tile_start = prm.tile_origin + (int64_t)tile * kTile;
    int stage_bytes = kTile + kHalo;
    if (tile_start + stage_bytes > prm.n) stage_bytes = (int)(prm.n - tile_start > 0 ? prm.n - tile_start : 0);
    {
      const int bulk = text_aligned ? (stage_bytes & ~15) : 0;
      for (int i = bulk + tid; i < stage_bytes; i += kThreads) sh.text[i] = prm.text[tile_start + i];
      mbar_wait(&sh.bar, (uint32_t)(it & 1));
      if (bulk != stage_bytes) __syncthreads();
    }
    TextView tv{prm.text, sh.text, tile_start, tile_start + stage_bytes};
    const int gen = it & 1;
    uint16_t *mine = sh.priv[gen] + tid * kPriv;
    // ---- scan (never waits on other CTAs), ranks inside the tile, publish the aggregate
    int cnt_a;
    bool ovf;
    const int cnt = dense_scan_chunk<false>(prm, sh, tv, tile_start, mine, cnt_a, ovf);
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    const int lane_prefix = inc - cnt;
    const int warp_total = __shfl_sync(0xffffffffu, inc, 31);
    const int overflow = __syncthreads_or(ovf);
    int total;
    const int wrank = block_exclusive_scan(lane == 0 ? warp_total : 0, sh.scan_ws, &total);
    const int rank = __shfl_sync(0xffffffffu, wrank, 0) + lane_prefix;
    if (tid == 0) lookback_publish(prm.desc, tile, (unsigned long long)total);
    if (overflow) {
      // recompute path (some thread found more than kPriv matches in its 144 bytes): flush what is
      // pending, resolve this tile now, walk the chunk again and emit directly while the text is staged
      if (pending) flush_pending();
      const unsigned long long b = lookback_resolve_block(prm.desc, tile, (unsigned long long)total, sh.lb_ws);
      if (tid == 0 && tile == prm.n_tiles - 1) *prm.total = b + (unsigned long long)total;
      unsigned long long o = b + (unsigned long long)rank;
      const int chunk_rel = tid * kChunk;
      const int valid = (int)min((int64_t)kTile, prm.n - tile_start);
      const int64_t chunk0 = tile_start + chunk_rel;
      uint32_t s2 = 0;
      int64_t q = chunk0 - (sh.hdr.max_len - 1);
      if (q < 0) q = 0;
      const int64_t stop = chunk0 < prm.n ? chunk0 : prm.n;
      for (; q < stop; ++q) s2 = *reinterpret_cast<const uint16_t *>(tbl + s2 + 2u * prm.text[q]);
      for (int i = 0; i < kChunk && chunk_rel + i < valid; ++i) {  // (the staged copy now holds state bytes)
        s2 = *reinterpret_cast<const uint16_t *>(tbl + s2 + 2u * prm.text[chunk0 + i]);
        if (s2 & 4u) {
          const uint32_t len = sh.hdr.len_lut[(s2 >> 3) & 31u];
          const int64_t start = chunk0 + i + 1 - (int64_t)len;
          if (start >= prm.key_lo && start < prm.key_hi) {
            if (o < (unsigned long long)prm.cap) st_span(prm.spans + 2 * o, start + prm.rebase, start + len + prm.rebase);
            ++o;
          }
        }
      }
      __syncthreads();
    }
    // ---- the text buffer is free (block scan ended with a barrier): prefetch the next tile, and while it
    //      is in flight flush the tile scanned one iteration ago
    if (tid == 0) take_ticket_and_load();
    if (pending) flush_pending();
    if (!overflow) {
      pending = true;
      p_tile = tile; p_cnt = cnt; p_cnt_a = cnt_a; p_rank = rank; p_total = total; p_gen = gen; p_start = tile_start;
    }
    __syncthreads();
    tile = sh.tile;
  }
  if (pending) flush_pending();
}

// ---------------------------------------------------------------------------------------------
// per-document search: match_next(doc, 0) for every document of a batch (the reference's
// callers loop over dfa.mojo:2241-2253).  One thread per document.  Pass 1 walks the document BACKWARDS
// through the reverse automaton (program.hpp: RevTable) -- one shared-memory lookup per byte, the same
// 4-instruction step as k_dense -- and remembers the last flagged position it sees, which is the leftmost
// position whose attempt succeeds: exactly the start the reference's ascending loop over candidates
// returns.  Pass 2 runs that one attempt forwards through the dense forward table for the longest accept.
// No candidate queue, no atomics, no divergence in pass 1.
//
// Fixed-size, 16-byte aligned documents are staged warp by warp: a warp owns 32 consecutive documents, reads
// them with coalesced 16-byte loads (8 lanes per document row) and parks them in rows of `line_pitch` bytes
// (an odd number of 16-byte units, so both the staging stores and each lane's LDS.128 of its own row hit 8
// distinct bank groups).  Warps are independent -- no block barrier, no mbarrier -- and persistent over
// batches so that the tables are loaded once.  (A TMA bulk copy per document was tried first: UBLKCP takes
// uniform operands, so the 32 copies of a warp are issued by a 32-iteration ELECT loop that cost 18 % of
// the kernel's instructions, profiles/r01_k_lines_ncu.txt.)  Ragged or unaligned documents are read from
// HBM byte by byte.
struct LinesShared {
  alignas(16) uint8_t tables[16];  // rev | fwd | staged documents follow
};

// the attempt at `best` (dfa.mojo:1979-2026): longest accept, stop at the first dead transition
template <class ByteAt>
__device__ __forceinline__ void lines_emit(const uint8_t *fwd, uint32_t fwd_start, ByteAt byte_at, int len, int best,
                                           int64_t *out) {
  int64_t s_rel = -1, e_rel = -1;
  if (best >= 0) {
    uint32_t s = fwd_start;
    int q = best, last = -1;
    while (q < len) {
      s = *reinterpret_cast<const uint16_t *>(fwd + s + 2u * byte_at(q));
      if (s == 0) break;
      ++q;
      if (s & 4u) last = q;
    }
    if (last >= 0) { s_rel = best; e_rel = last; }
  }
  st_span(out, s_rel, e_rel);
}

__global__ void __launch_bounds__(kThreads) k_lines(const LinesParams prm) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  LinesShared &sh = *reinterpret_cast<LinesShared *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  {
    const uint4 *s4 = reinterpret_cast<const uint4 *>(prm.tables);
    uint4 *d4 = reinterpret_cast<uint4 *>(sh.tables);
    for (int i = tid; i < (prm.rev_bytes + prm.fwd_bytes) / 16; i += kThreads) d4[i] = s4[i];
  }
  __syncthreads();
  const uint8_t *rev = sh.tables, *fwd = sh.tables + prm.rev_bytes;
  constexpr int kWarpsPerCta = kThreads / 32;

  if (prm.staged) {
    const int pitch = prm.line_pitch, stride = (int)prm.doc_stride, nv = stride >> 4;
    uint8_t *rows = sh.tables + prm.rev_bytes + prm.fwd_bytes + (size_t)warp * 32 * pitch;  // this warp's 32 rows
    const uint8_t *mine = rows + (size_t)lane * pitch;
    const int64_t n_groups = (prm.n_docs + 31) / 32;  // a group = 32 consecutive documents = one warp batch
    for (int64_t grp = (int64_t)blockIdx.x * kWarpsPerCta + warp; grp < n_groups; grp += (int64_t)gridDim.x * kWarpsPerCta) {
      const int64_t d0 = grp * 32;
      const int in_grp = (int)min((int64_t)32, prm.n_docs - d0);
      // ---- stage: the group's documents are contiguous in HBM; chunk c (16 bytes) of that range belongs to
      //      row c / nv, column c % nv
      const uint4 *src = reinterpret_cast<const uint4 *>(prm.text + d0 * stride);
      const int chunks = in_grp * nv;
      __syncwarp();  // the previous batch's rows are no longer read
      for (int c0 = 0; c0 < chunks; c0 += 128) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = c0 + 32 * u + lane;
          if (c < chunks) v[u] = __ldg(src + c);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = c0 + 32 * u + lane;
          if (c < chunks) *reinterpret_cast<uint4 *>(rows + (size_t)(c / nv) * pitch + 16 * (c % nv)) = v[u];
        }
      }
      __syncwarp();
      if (lane < in_grp) {
        uint32_t r = 0, best_pk = 0;
        int best_word = -1;  // index of the 4-byte word that holds the leftmost flagged byte
#pragma unroll 2
        for (int k = nv - 1; k >= 0; --k) {
          const uint4 w4 = *reinterpret_cast<const uint4 *>(mine + 16 * k);
          const uint32_t ws[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
          for (int wi = 3; wi >= 0; --wi) {
            uint32_t pk = 0;
#pragma unroll
            for (int j = 3; j >= 0; --j) {
              const uint32_t b = __byte_perm(ws[wi], 0, 0x4440 + j);
              uint32_t addr;
              asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(addr) : "r"(b), "r"(r));
              r = *reinterpret_cast<const uint16_t *>(rev + addr);
              pk = __byte_perm(pk, r, j == 0 ? 0x3214 : j == 1 ? 0x3240 : j == 2 ? 0x3410 : 0x4210);
            }
            if (pk & 0x04040404u) { best_pk = pk; best_word = 4 * k + wi; }
          }
        }
        const int best = best_word >= 0 ? 4 * best_word + ((__ffs(best_pk & 0x04040404u) - 1) >> 3) : -1;
        lines_emit(fwd, (uint32_t)prm.fwd_start, [&](int q) { return (uint32_t)mine[q]; }, stride, best,
                   prm.out + 2 * (d0 + lane));
      }
    }
    return;
  }

  const int64_t n_batches = (prm.n_docs + kThreads - 1) / kThreads;
  for (int64_t batch = blockIdx.x; batch < n_batches; batch += gridDim.x) {
    const int64_t d = batch * kThreads + tid;
    if (d >= prm.n_docs) break;
    const int64_t lo = prm.doc_offsets ? prm.doc_offsets[d] : d * prm.doc_stride;
    const int64_t hi = prm.doc_offsets ? prm.doc_offsets[d + 1] : min(lo + prm.doc_stride, prm.n);
    const uint8_t *doc = prm.text + lo;
    const int len = (int)min(hi - lo, (int64_t)0x7fffffff);
    int best = -1;
    uint32_t r = 0;
    for (int q = len - 1; q >= 0; --q) {
      r = *reinterpret_cast<const uint16_t *>(rev + r + 2u * doc[q]);
      if (r & 4u) best = q;
    }
    lines_emit(fwd, (uint32_t)prm.fwd_start, [&](int q) { return (uint32_t)doc[q]; }, len, best, prm.out + 2 * d);
  }
}

// Batch search with a required byte (RevTable::req_byte, e.g. '@' for the e-mail pattern): a document that
// does not hold that byte cannot match, so only the others are walked.  One CTA per SM, kReqWarps
// independent warps.  A warp loads 32 consecutive documents with coalesced 16-byte loads and tests every
// 16-byte piece for the byte while it is still in registers (SWAR zero-byte test on piece ^ byte x 4); the
// OR over a document's pieces says whether it is a candidate.  Documents that are not get (-1, -1) at once and
// never touch shared memory; the pieces of candidates go to the next free rows of a 64-row ring (same bank-
// skewed row pitch as k_lines) together with the document index.  As soon as 32 rows are pending, every
// lane walks one of them exactly as k_lines does (reverse automaton for the leftmost start, forward table for
// the longest accept), so the walk always runs with full warps however sparse the candidates are.  The ring
// holds < 32 pending rows before a batch and receives <= 32 new ones: 64 rows never overflow.
constexpr int kReqWarps = 16;  // 512 threads x 128 registers: two batches of pieces live in registers
constexpr int kReqThreads = 32 * kReqWarps;
constexpr int kReqRing = 64;
constexpr int kReqMaxVecs = 8;  // documents of up to 128 bytes: a lane holds one 16-byte piece per vector slot

__device__ __forceinline__ void lines_walk_row(const uint8_t *rev, const uint8_t *fwd, uint32_t fwd_start,
                                               const uint8_t *mine, int nv, int stride, int64_t *out) {
  uint32_t r = 0, best_pk = 0;
  int best_word = -1;  // index of the 4-byte word that holds the leftmost flagged byte
#pragma unroll 2
  for (int k = nv - 1; k >= 0; --k) {
    const uint4 w4 = *reinterpret_cast<const uint4 *>(mine + 16 * k);
    const uint32_t ws[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
    for (int wi = 3; wi >= 0; --wi) {
      uint32_t pk = 0;
#pragma unroll
      for (int j = 3; j >= 0; --j) {
        const uint32_t b = __byte_perm(ws[wi], 0, 0x4440 + j);
        uint32_t addr;
        asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(addr) : "r"(b), "r"(r));
        r = *reinterpret_cast<const uint16_t *>(rev + addr);
        pk = __byte_perm(pk, r, j == 0 ? 0x3214 : j == 1 ? 0x3240 : j == 2 ? 0x3410 : 0x4210);
      }
      if (pk & 0x04040404u) { best_pk = pk; best_word = 4 * k + wi; }
    }
  }
  const int best = best_word >= 0 ? 4 * best_word + ((__ffs(best_pk & 0x04040404u) - 1) >> 3) : -1;
  lines_emit(fwd, fwd_start, [&](int q) { return (uint32_t)mine[q]; }, stride, best, out);
}

template <int NV>  // NV = 16-byte pieces per document when known at compile time (8: 128-byte lines), 0 = runtime
__global__ void __launch_bounds__(kReqThreads, 1) k_lines_req(const LinesParams prm) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  LinesShared &sh = *reinterpret_cast<LinesShared *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  {
    const uint4 *s4 = reinterpret_cast<const uint4 *>(prm.tables);
    uint4 *d4 = reinterpret_cast<uint4 *>(sh.tables);
    for (int i = tid; i < (prm.rev_bytes + prm.fwd_bytes) / 16; i += kReqThreads) d4[i] = s4[i];
  }
  __syncthreads();
  const uint8_t *rev = sh.tables, *fwd = sh.tables + prm.rev_bytes;
  const int pitch = prm.line_pitch, stride = (int)prm.doc_stride, nv = NV ? NV : (stride >> 4);
  const int nv_shift = (nv & (nv - 1)) == 0 ? 31 - __clz(nv) : -1;  // power of two: shifts instead of divisions
  const size_t per_warp = (size_t)kReqRing * pitch + (size_t)kReqRing * sizeof(int64_t);
  uint8_t *rows = sh.tables + prm.rev_bytes + prm.fwd_bytes + (size_t)warp * per_warp;
  int64_t *ids = reinterpret_cast<int64_t *>(rows + (size_t)kReqRing * pitch);
  const uint32_t rq = (uint32_t)prm.req_byte * 0x01010101u;
  const uint32_t below = (1u << lane) - 1u;
  uint32_t head = 0, tail = 0;  // ring counters of this warp (warp-uniform); row = counter & 63
  const int64_t n_groups = (prm.n_docs + 31) / 32;
  const int64_t gstep = (int64_t)gridDim.x * kReqWarps;
  // the pieces of batch g: piece c = 32 j + lane of the batch's contiguous byte range, j < nv
  auto load_batch = [&](int64_t g, uint4 (&dst)[kReqMaxVecs]) {
    const int64_t d0 = g * 32;
    const int chunks = (int)min((int64_t)32, prm.n_docs - d0) * nv;
    const uint4 *src = reinterpret_cast<const uint4 *>(prm.text + d0 * stride);
#pragma unroll
    for (int j = 0; j < kReqMaxVecs; ++j) {
      const int c = 32 * j + lane;
      if (j < nv && c < chunks) dst[j] = __ldg(src + c);
    }
  };
  // batches by ticket, one per warp at a time, fetched two batches ahead (the atomic's latency hides behind a batch).
  // A static split made the kernel as slow as its slowest CTA: on a multi-GPU job the previous step's all-gather
  // kernel can still sit on an SM when this kernel starts, and the CTA of that SM began late (C3 at N = 8: 0.58 ms
  // against 0.49 alone).  Without a ticket counter (prm.ticket == nullptr) the split is static.
  auto take = [&](int64_t after) -> int64_t {
    if (!prm.ticket) return after + gstep;
    unsigned int t = 0;
    if (lane == 0) t = atomicAdd(prm.ticket, 1u);
    return (int64_t)__shfl_sync(0xffffffffu, t, 0);
  };
  int64_t grp = prm.ticket ? take(0) : (int64_t)blockIdx.x * kReqWarps + warp;
  int64_t nxt = take(grp);
  uint4 v[kReqMaxVecs];
  if (grp < n_groups) load_batch(grp, v);
  for (; grp < n_groups;) {
    // the next batch's loads are in flight while this one is filtered, parked and (now and then) walked
    uint4 nx[kReqMaxVecs];
    if (nxt < n_groups) load_batch(nxt, nx);
    const int64_t after_next = take(nxt);
    const int64_t d0 = grp * 32;
    const int in_grp = (int)min((int64_t)32, prm.n_docs - d0);
    const int chunks = in_grp * nv;
    uint32_t fl = 0;  // bit d: one of my pieces of document d0 + d holds the byte
#pragma unroll
    for (int j = 0; j < kReqMaxVecs; ++j) {
      const int c = 32 * j + lane;
      if (j < nv && c < chunks) {
        const uint32_t x0 = v[j].x ^ rq, x1 = v[j].y ^ rq, x2 = v[j].z ^ rq, x3 = v[j].w ^ rq;
        const uint32_t z = ((x0 - 0x01010101u) & ~x0) | ((x1 - 0x01010101u) & ~x1) | ((x2 - 0x01010101u) & ~x2) |
                           ((x3 - 0x01010101u) & ~x3);
        const int d = nv_shift >= 0 ? (c >> nv_shift) : (c / nv);
        if (z & 0x80808080u) fl |= 1u << d;
      }
    }
    const uint32_t m = __reduce_or_sync(0xffffffffu, fl);  // candidate documents of this batch
    if (lane < in_grp && !((m >> lane) & 1u)) st_span(prm.out + 2 * (d0 + lane), -1, -1);
#pragma unroll
    for (int j = 0; j < kReqMaxVecs; ++j) {
      const int c = 32 * j + lane;
      if (j < nv && c < chunks) {
        const int d = nv_shift >= 0 ? (c >> nv_shift) : (c / nv);
        if ((m >> d) & 1u) {
          const uint32_t slot = (tail + (uint32_t)__popc(m & ((1u << d) - 1u))) & (kReqRing - 1);
          *reinterpret_cast<uint4 *>(rows + (size_t)slot * pitch + 16 * (c - d * nv)) = v[j];
        }
      }
    }
    if ((m >> lane) & 1u) ids[(tail + (uint32_t)__popc(m & below)) & (kReqRing - 1)] = d0 + lane;
    tail += (uint32_t)__popc(m);
    __syncwarp();
    if (tail - head >= 32u) {
      const uint32_t slot = (head + (uint32_t)lane) & (kReqRing - 1);
      lines_walk_row(rev, fwd, (uint32_t)prm.fwd_start, rows + (size_t)slot * pitch, nv, stride, prm.out + 2 * ids[slot]);
      head += 32u;
      __syncwarp();  // the walked rows may be overwritten by the next batch
    }
#pragma unroll
    for (int j = 0; j < kReqMaxVecs; ++j) v[j] = nx[j];
    grp = nxt;
    nxt = after_next;
  }
  if ((uint32_t)lane < tail - head) {
    const uint32_t slot = (head + (uint32_t)lane) & (kReqRing - 1);
    lines_walk_row(rev, fwd, (uint32_t)prm.fwd_start, rows + (size_t)slot * pitch, nv, stride, prm.out + 2 * ids[slot]);
  }
}

// ---------------------------------------------------------------------------------------------
// class runs
constexpr int kRunTok = 12;  // matches a thread buffers before it has to walk its chunk a second time
struct RunsShared {
  DevProg prog;
  alignas(16) uint8_t text[kTile + kHalo];
  alignas(8) uint64_t bar;
  int scan_ws[kThreads / 32 + 2];
  int tile;
  unsigned long long base;
  long long warp_len[kThreads / 32];
  int warp_all[kThreads / 32];
  uint16_t tok[kThreads * kRunTok];  // UNIFORM_RUN: (start in chunk) | (length << 8) of a thread's matches
  uint32_t head_bits[kThreads + 1];  // UNIFORM_RUN: membership bits of the first 32 bytes of every chunk (+ the halo's)
};

// membership bits of 16 bytes
template <int FILT>
__device__ __forceinline__ uint32_t class_bits16(const uint4 &w, const Filter &f, const uint8_t *cand) {
  uint32_t m = 0;
  const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t fl;
    if (FILT == FILT_NEEDLES) {
      fl = swar_eq(ws[i], f.n0);
      if (f.n_needles > 1) fl |= swar_eq(ws[i], f.n1);
      if (f.n_needles > 2) fl |= swar_eq(ws[i], f.n2);
      if (f.n_needles > 3) fl |= swar_eq(ws[i], f.n3);
    } else {
      fl = filter_word<FILT>(ws[i], f, cand);
    }
    m |= flags_to_nibble(fl) << (4 * i);
  }
  return m;
}

// 144 membership bits of one thread's chunk
struct ChunkBits {
  uint32_t w[5];  // bit i of w[i/32] = byte i of the chunk (bits 144..159 of w[4] are zero)
  // first position >= i whose bit equals `set` (kChunk if none); fully unrolled so w[] stays in registers
  __device__ __forceinline__ int find(int i, bool set) const {
    int res = kChunk;
#pragma unroll
    for (int k = 4; k >= 0; --k) {
      uint32_t x = set ? w[k] : ~w[k];
      if (k == 4) x &= 0xffffu;
      if (32 * k + 31 < i) x = 0;
      else if (32 * k < i) x &= ~0u << (i - 32 * k);
      if (x) res = 32 * k + __ffs(x) - 1;
    }
    return res;
  }
};

template <int FILT>
__global__ void __launch_bounds__(kThreads, 3) k_runs(const RunsParams prm) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  RunsShared &sh = *reinterpret_cast<RunsShared *>(smem_raw);
  const int tid = threadIdx.x;
  if (tid == 0) sh.tile = prm.tile_base + (int)atomicAdd(prm.ticket, 1u);
  load_prog(&sh.prog, prm.prog);
  __syncthreads();
  const int tile = sh.tile;
  const int64_t tile_start = prm.tile_origin + (int64_t)tile * kTile;
  int stage_bytes = kTile + kHalo;
  if (tile_start + stage_bytes > prm.n) stage_bytes = (int)(prm.n - tile_start > 0 ? prm.n - tile_start : 0);
  stage_tile(sh.text, prm.text, tile_start, stage_bytes, &sh.bar);
  const uint8_t *lut = sh.prog.cand;
  TextView tv{prm.text, sh.text, tile_start, tile_start + stage_bytes};

  const int64_t chunk0 = tile_start + (int64_t)tid * kChunk;
  // membership bits.  CLASS_RUNS treats the key window as the whole text (runs are clipped to it;
  // shards are stitched by the caller); UNIFORM_RUN needs the real run structure of the buffer.
  const int64_t vis_lo = prm.mode == MRX_MODE_CLASS_RUNS ? prm.key_lo : 0;
  const int64_t vis_hi = prm.mode == MRX_MODE_CLASS_RUNS ? prm.key_hi : prm.n;
  ChunkBits cb;
#pragma unroll
  for (int i = 0; i < 5; ++i) cb.w[i] = 0;
#pragma unroll
  for (int v = 0; v < kVecs; ++v) {
    const int64_t pos0 = chunk0 + 16 * v;
    uint32_t m = 0;
    if (pos0 < vis_hi && pos0 + 16 > vis_lo) {
      const uint4 w = *reinterpret_cast<const uint4 *>(sh.text + tid * kChunk + 16 * v);
      m = class_bits16<FILT>(w, prm.filt, lut);
      if (pos0 + 16 > vis_hi) m &= (1u << (int)(vis_hi - pos0)) - 1u;
      if (pos0 < vis_lo) m &= ~((1u << (int)(vis_lo - pos0)) - 1u);
    }
    cb.w[v >> 1] |= m << (16 * (v & 1));
  }
  const bool prev_in = chunk0 > vis_lo && chunk0 <= vis_hi && lut[tv.at(chunk0 - 1)];
  const bool next_in = (chunk0 + kChunk < vis_hi) && chunk0 + kChunk >= vis_lo && lut[tv.at(chunk0 + kChunk)];

  if (prm.mode == MRX_MODE_CLASS_RUNS) {
    // starts: in-class byte whose predecessor is not; ends: in-class byte whose successor is not.
    // The k-th start pairs with the k-th end globally, so starts and ends are ranked independently
    // and #ends-before-tile = #starts-before-tile - (a run is open across the tile start).
    uint32_t st[5], en[5];
    int ns = 0, ne = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const uint32_t carry_in = (i == 0) ? (prev_in ? 1u : 0u) : (cb.w[i - 1] >> 31);
      const uint32_t nxt = (i == 4) ? (next_in ? (1u << 16) : 0u) : 0u;
      const uint32_t up = (i < 4) ? (cb.w[i + 1] << 31) : 0u;
      st[i] = cb.w[i] & ~((cb.w[i] << 1) | carry_in);
      en[i] = cb.w[i] & ~((cb.w[i] >> 1) | up | (i == 4 ? (nxt >> 1) : 0u));
      ns += __popc(st[i]);
      ne += __popc(en[i]);
    }
    int total_s;
    const int rank_s = block_exclusive_scan(ns, sh.scan_ws, &total_s);
    int total_e;
    const int rank_e = block_exclusive_scan(ne, sh.scan_ws, &total_e);
    if (tid < 32) {
      const unsigned long long b = lookback_exclusive(prm.desc, tile, (unsigned long long)total_s);
      if (tid == 0) {
        sh.base = b;
        if (tile == prm.n_tiles - 1) *prm.total = b + (unsigned long long)total_s;
      }
    }
    __syncthreads();
    const bool open_at_tile = tile_start > vis_lo && tile_start < vis_hi && lut[tv.at(tile_start - 1)] && lut[tv.at(tile_start)];
    unsigned long long os = sh.base + rank_s;
    unsigned long long oe = sh.base - (open_at_tile ? 1ull : 0ull) + rank_e;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      uint32_t f = st[i];
      while (f) {
        const int b = __ffs(f) - 1;
        f &= f - 1;
        if (os < (unsigned long long)prm.cap) prm.spans[2 * os] = chunk0 + 32 * i + b + prm.rebase;
        ++os;
      }
      f = en[i];
      while (f) {
        const int b = __ffs(f) - 1;
        f &= f - 1;
        if (oe < (unsigned long long)prm.cap) prm.spans[2 * oe + 1] = chunk0 + 32 * i + b + 1 + prm.rebase;
        ++oe;
      }
    }
    return;
  }

  // ---- UNIFORM_RUN: chain automaton over one class, accepting depths [mn, K] -------------------
  // distance from the run start at the first byte of this chunk
  int lead = 0;
  {
    bool all = true;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const uint32_t full = (i == 4) ? 0xffffu : 0xffffffffu;
      if (all) {
        if (cb.w[i] == full) lead += (i == 4) ? 16 : 32;
        else { lead += __ffs(~cb.w[i]) - 1; all = false; }
      }
    }
  }
  const bool all_in = lead == kChunk;
  int trail = 0;
  if (!all_in) {
    bool go = true;
#pragma unroll
    for (int i = 4; i >= 0; --i) {
      const int bits = (i == 4) ? 16 : 32;
      const uint32_t full = (i == 4) ? 0xffffu : 0xffffffffu;
      if (go) {
        if (cb.w[i] == full) trail += bits;
        else { trail += __clz(~(cb.w[i] << (32 - bits))); go = false; }
      }
    }
  }
  // inclusive scan of (all, len) with (a1,l1)*(a2,l2) = (a1&a2, a2 ? l1+l2 : l2)
  long long len = all_in ? kChunk : trail;
  int allf = all_in ? 1 : 0;
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long pl = __shfl_up_sync(0xffffffffu, len, o);
    const int pa = __shfl_up_sync(0xffffffffu, allf, o);
    if (lane >= o) {
      if (allf) len += pl;
      allf &= pa;
    }
  }
  if (lane == 31) { sh.warp_len[warp] = len; sh.warp_all[warp] = allf; }
  // the in-class bytes that FOLLOW a chunk are the head bits of the next one (the last chunk looks into the halo)
  sh.head_bits[tid] = cb.w[0];
  if (tid == kThreads - 1) {
    uint32_t hb = 0;
    const int64_t h0 = tile_start + kTile;
    if (h0 < prm.n && stage_bytes >= kTile + 32) {
      const uint4 a = *reinterpret_cast<const uint4 *>(sh.text + kTile), b = *reinterpret_cast<const uint4 *>(sh.text + kTile + 16);
      hb = class_bits16<FILT>(a, prm.filt, lut) | (class_bits16<FILT>(b, prm.filt, lut) << 16);
    } else {
      for (int i = 0; i < 32 && h0 + i < prm.n; ++i) hb |= (lut[tv.at(h0 + i)] ? 1u : 0u) << i;
    }
    if (h0 + 32 > prm.n) hb &= h0 < prm.n ? ((1u << (int)(prm.n - h0)) - 1u) : 0u;
    sh.head_bits[kThreads] = hb;
  }
  __syncthreads();
  // carry entering this thread = combine(tile carry, warps before, lanes before)
  long long carry = prm.tile_carry ? prm.tile_carry[tile] : 0;  // distance at tile start (0 if not in a run)
  for (int w = 0; w < warp; ++w) carry = sh.warp_all[w] ? carry + sh.warp_len[w] : sh.warp_len[w];
  {
    const long long pl = __shfl_up_sync(0xffffffffu, len, 1);
    const int pa = __shfl_up_sync(0xffffffffu, allf, 1);
    if (lane > 0) carry = pa ? carry + pl : pl;
  }
  // `carry` = number of consecutive in-class bytes immediately before this chunk
  const int K = prm.run_max, MN = prm.run_min;

  int cnt = 0;
  unsigned long long out_idx = 0;
  uint16_t *tok = sh.tok + tid * kRunTok;
  const bool tokens_ok = K <= 255;  // a match length fits the token
  const uint32_t next_head = sh.head_bits[tid + 1];
  for (int pass = 0; pass < 2; ++pass) {
    cnt = 0;
    int i = 0;
    while (i < kChunk) {
      // find next run segment [i0, i1) inside the chunk
      const int i0 = cb.find(i, true);
      if (i0 >= kChunk) break;
      const int i1 = cb.find(i0, false);
      const long long d0 = (i0 == 0) ? carry : 0;  // distance of byte i0 from its run start
      int ext = 0;                                 // in-class bytes following the chunk, capped at K
      if (i1 == kChunk) {
        if (K <= 32) {
          const int run = __ffs(~next_head) - 1;  // leading in-class bytes of the next chunk (32 if all)
          ext = run < 0 ? 32 : run;
          if (ext > K) ext = K;
        } else {
          int64_t q = chunk0 + kChunk;
          while (ext < K && q < prm.n && lut[tv.at(q)]) { ++ext; ++q; }
        }
      }
      int first = i0;  // a run that starts inside the chunk starts a match; one that continues from before
      if (d0 != 0) {   // the chunk start resumes at its phase (64-bit modulo only for runs longer than 2 GiB)
        const int ph = d0 < 0x7fffffffll ? (int)((unsigned)d0 % (unsigned)K) : (int)(d0 % K);
        first = i0 + (ph ? K - ph : 0);
      }
      for (int s = first; s < i1; s += K) {
        const int rem = (i1 - s) + ext;  // exact when < K
        const int L = rem >= K ? K : rem;
        if (L < MN) continue;
        const int64_t pos = chunk0 + s;
        if (pos < prm.key_lo || pos >= prm.key_hi) continue;
        if (pass == 1) {
          const unsigned long long o = out_idx + cnt;
          if (o < (unsigned long long)prm.cap) st_span(prm.spans + 2 * o, pos + prm.rebase, pos + L + prm.rebase);
        } else if (cnt < kRunTok) {
          tok[cnt] = (uint16_t)(s | (L << 8));
        }
        ++cnt;
      }
      i = i1;
    }
    if (pass == 0) {
      int total;
      const int rank = block_exclusive_scan(cnt, sh.scan_ws, &total);
      if (tid < 32) {
        const unsigned long long b = lookback_exclusive(prm.desc, tile, (unsigned long long)total);
        if (tid == 0) {
          sh.base = b;
          if (tile == prm.n_tiles - 1) *prm.total = b + (unsigned long long)total;
        }
      }
      __syncthreads();
      out_idx = sh.base + (unsigned long long)rank;
      if (tokens_ok && cnt <= kRunTok) {  // the usual case: emit from the buffered tokens, no second walk
        for (int k = 0; k < cnt; ++k) {
          const unsigned long long o = out_idx + k;
          const int64_t pos = chunk0 + (tok[k] & 0xff);
          if (o < (unsigned long long)prm.cap) st_span(prm.spans + 2 * o, pos + prm.rebase, pos + (tok[k] >> 8) + prm.rebase);
        }
        return;
      }
    }
  }
}

// trailing in-class run length of every tile (and whether the whole tile is in class)
__global__ void k_tile_trailing(const uint8_t *text, int64_t n, int64_t tile_origin, int n_tiles, const DevProg *prog,
                                long long *trail, int *all) {
  const int tile = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  if (tile >= n_tiles) return;
  const int lane = threadIdx.x & 31;
  const int64_t begin = tile_origin + (int64_t)tile * kTile;
  int64_t end = begin + kTile;
  if (end > n) end = n;
  // walk backwards 32 bytes at a time
  long long t = 0;
  bool stop = false;
  for (int64_t hi = end; hi > begin && !stop; hi -= 32) {
    const int64_t q = hi - 1 - lane;
    const bool in = (q >= begin) && prog->cand[text[q]];
    const unsigned m = __ballot_sync(0xffffffffu, in);
    const int width = (int)min((int64_t)32, hi - begin);
    const unsigned fullm = width == 32 ? 0xffffffffu : ((1u << width) - 1u);
    if ((m & fullm) == fullm) t += width;
    else { t += __ffs(~m) - 1; stop = true; }
  }
  if (lane == 0) {
    trail[tile] = t;
    all[tile] = (!stop && end - begin == kTile) ? 1 : 0;
    if (!stop && end - begin < kTile) all[tile] = 0;
  }
}

// carry[tile] = in-class bytes immediately before the tile start (sequential over tiles; n_tiles is
// text/36 KB, so this is microseconds)
// carry[t] = length of the class run that ends exactly at the start of tile t.  Segmented scan over
// (all, trail) pairs with  (a1,l1) o (a2,l2) = (a1 & a2, a2 ? l1 + l2 : l2):  one block of 1024 threads,
// every thread owns a contiguous slice of tiles (serial), slices are combined by a block scan.
__global__ void __launch_bounds__(1024) k_carry_scan(int n_tiles, const long long *trail, const int *all, long long *carry,
                                                     const long long *carry0) {
  __shared__ long long s_len[1024];
  __shared__ int s_all[1024];
  const int tid = threadIdx.x;
  const int per = (n_tiles + 1023) / 1024;
  const int t0 = min(tid * per, n_tiles), t1 = min(t0 + per, n_tiles);
  long long len = 0;
  int a = 1;
  for (int t = t0; t < t1; ++t) {
    const int at = all[t];
    len = at ? len + trail[t] : trail[t];
    a &= at;
  }
  s_len[tid] = len;
  s_all[tid] = a;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {  // inclusive Hillis-Steele scan of the slice summaries
    long long l2 = s_len[tid];
    int a2 = s_all[tid];
    if (tid >= o) {
      const long long l1 = s_len[tid - o];
      const int a1 = s_all[tid - o];
      if (a2) l2 += l1;
      a2 &= a1;
    }
    __syncthreads();
    s_len[tid] = l2;
    s_all[tid] = a2;
    __syncthreads();
  }
  // run length entering this thread's slice; the run that ends right before the first tile (carry0, a windowed
  // text) continues into it as long as every tile so far is all in class
  long long c = tid > 0 ? s_len[tid - 1] : 0;
  if (carry0 && (tid == 0 || s_all[tid - 1])) c += *carry0;
  for (int t = t0; t < t1; ++t) {
    carry[t] = c;
    c = all[t] ? c + trail[t] : trail[t];
  }
}

// length of the class run that ends right before `origin` (UNIFORM_RUN on a windowed text: the distance of the
// first tile's first byte from its run start).  One block walks backwards, kThreads bytes per round.
__global__ void __launch_bounds__(kThreads) k_back_run(const uint8_t *text, int64_t origin, const DevProg *prog, long long *out) {
  __shared__ int s_first;
  long long len = 0;
  for (int64_t hi = origin; hi > 0; hi -= kThreads) {
    if (threadIdx.x == 0) s_first = kThreads;
    __syncthreads();
    const int64_t q = hi - 1 - threadIdx.x;
    const bool in = q >= 0 && prog->cand[text[q]];
    if (!in) atomicMin(&s_first, (int)threadIdx.x);
    __syncthreads();
    const int first = s_first;
    __syncthreads();
    len += first;
    if (first < kThreads) break;
  }
  if (threadIdx.x == 0) *out = len;
}

// ---------------------------------------------------------------------------------------------
// search helpers
__global__ void k_attempt_one(const ScanParams prm, int64_t pos, int64_t *out) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  Attempt at;
  at.p = prm.prog;
  at.tv.g = prm.text;
  at.tv.s = nullptr;
  at.tv.s_begin = at.tv.s_end = 0;
  at.lo = 0;
  at.n = prm.n;
  Hit h;
  if (try_candidate(prm, at, pos, h)) { out[0] = h.s; out[1] = h.e; out[2] = 1; }
  else { out[0] = out[1] = -1; out[2] = 0; }
}

// batch search for plans that need no scan: `.*` (every document matches whole, matcher.mojo:762-766) and
// `^`-anchored patterns (one attempt at the document start, dfa.mojo:1888-1892).  One thread per document.
__global__ void k_batch_trivial(const ScanParams prm, int wildcard, int64_t *out) {
  const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= prm.n_docs) return;
  const int64_t lo = prm.doc_offsets ? prm.doc_offsets[d] : d * prm.doc_stride;
  const int64_t hi = prm.doc_offsets ? prm.doc_offsets[d + 1] : min(lo + prm.doc_stride, prm.n);
  int64_t s_rel = -1, e_rel = -1;
  if (wildcard) {
    s_rel = 0;
    e_rel = hi - lo;
  } else {
    Attempt at;
    at.p = prm.prog;
    at.tv.g = prm.text;
    at.tv.s = nullptr;
    at.tv.s_begin = at.tv.s_end = 0;
    at.lo = lo;
    at.n = hi;
    Hit h;
    if (try_candidate(prm, at, lo, h)) { s_rel = h.s - lo; e_rel = h.e - lo; }
  }
  st_span(out + 2 * d, s_rel, e_rel);
}

__global__ void k_batch_finalize(const unsigned long long *doc_best, int64_t n_docs, const int64_t *doc_offsets,
                                 int64_t doc_stride, int64_t n, int empty_at_end, int64_t *out) {
  const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n_docs) return;
  const unsigned long long v = doc_best[d];
  if (v == ~0ull) {
    if (empty_at_end) {
      const int64_t lo = doc_offsets ? doc_offsets[d] : d * doc_stride;
      const int64_t hi = doc_offsets ? doc_offsets[d + 1] : min(lo + doc_stride, n);
      out[2 * d] = hi - lo; out[2 * d + 1] = hi - lo;
    } else {
      out[2 * d] = -1; out[2 * d + 1] = -1;
    }
  } else {
    out[2 * d] = (int64_t)(v >> 32);
    out[2 * d + 1] = (int64_t)(v & 0xffffffffull);
  }
}

__global__ void k_fill_u64(unsigned long long *p, int64_t n, unsigned long long v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
// p[i] += v (piece offsets of one chunk of a chunked split -> offsets in the whole output)
__global__ void k_add_i64(int64_t *p, int64_t n, int64_t v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] += v;
}

// ---------------------------------------------------------------------------------------------
// greedy selection over a potential list sorted by key (= the reference's `pos = end` restart
// loops, dfa.mojo:2078-2128 / matcher.mojo:876-897, applied to precomputed attempts).
//   sync node: key >= every earlier end  => no earlier match can cover it => always visited.
//   clusters between sync nodes are resolved by one thread each (typically 1-3 entries).
__global__ void k_select_prefix_max(const int64_t *spans, int64_t m, unsigned long long *desc, unsigned int *ticket,
                                    int64_t *pmax /* exclusive prefix max of ends */) {
  __shared__ int s_tile;
  __shared__ long long s_warp[kThreads / 32];
  __shared__ long long s_base;
  if (threadIdx.x == 0) s_tile = (int)atomicAdd(ticket, 1u);
  __syncthreads();
  const int tile = s_tile;
  const int64_t i = (int64_t)tile * kThreads + threadIdx.x;
  const long long NEG = -1;
  long long v = (i < m) ? (long long)spans[2 * i + 1] : NEG;
  long long inc = v;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc = max(inc, t);
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  long long before = NEG;
  for (int w = 0; w < warp; ++w) before = max(before, s_warp[w]);
  long long block_max = NEG;
  for (int w = 0; w < kThreads / 32; ++w) block_max = max(block_max, s_warp[w]);
  // look-back with max: reuse the sum look-back on (value+1) is not valid for max; do a max look-back
  if (threadIdx.x < 32) {
    unsigned long long agg = (unsigned long long)(block_max + 1);  // >= 0
    long long excl = 0;
    if (tile == 0) {
      if (lane == 0) st_volatile_u64(desc, kFlagPre | agg);
    } else {
      if (lane == 0) st_volatile_u64(desc + tile, kFlagAgg | agg);
      int idx = tile - 1;
      unsigned long long run = 0;
      while (true) {
        const int my = idx - lane;
        unsigned long long d = kFlagPre;
        if (my >= 0) { do { d = ld_volatile_u64(desc + my); } while ((d >> 62) == 0); }
        const unsigned pre_mask = __ballot_sync(0xffffffffu, (d >> 62) == 2);
        const int first_pre = pre_mask ? (__ffs(pre_mask) - 1) : 32;
        unsigned long long x = (lane <= first_pre) ? (d & kValMask) : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x = max(x, __shfl_xor_sync(0xffffffffu, x, o));
        run = max(run, x);
        if (pre_mask) break;
        idx -= 32;
      }
      excl = (long long)run;
      if (lane == 0) st_volatile_u64(desc + tile, kFlagPre | max(run, agg));
    }
    if (lane == 0) s_base = excl - 1;  // back to the "end" domain (-1 = none)
  }
  __syncthreads();
  long long ex = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane == 0) ex = NEG;
  ex = max(ex, before);
  ex = max(ex, s_base);
  if (i < m) pmax[i] = ex;
}

__global__ void k_select_walk(const int64_t *spans, const int64_t *keys, const int64_t *pmax, int64_t m, uint8_t *sel) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int64_t key_i = keys ? keys[i] : spans[2 * i];
  if (key_i < pmax[i]) return;  // not a sync node: some earlier potential match may cover it
  // resolve the cluster that starts here
  int64_t cover = key_i;
  for (int64_t j = i; j < m; ++j) {
    const int64_t kj = keys ? keys[j] : spans[2 * j];
    if (j > i && kj >= pmax[j]) break;  // next sync node
    if (kj >= cover) {
      sel[j] = 1;
      const int64_t s = spans[2 * j], e = spans[2 * j + 1];
      cover = (e > kj && e > s) ? e : kj + 1;  // `pos = end`, or `pos += 1` for an empty match
      if (e <= kj) cover = kj + 1;
    } else {
      sel[j] = 0;
    }
  }
}

// The restart chain of match_next behind a MemchrPrefilter (matcher.mojo:784-796 inside the loops of sub / findall
// callers): pos -> c = next occurrence of the literal at or after pos -> the first potential match starting at or
// after c -> pos = its end (one past its start for an empty match).  One thread walks the two sorted lists; this path
// serves patterns whose prefilter literal is not a prefix of the match, a corner of the reference's behaviour
// (SURVEY appendix B-10), not a throughput path.
__global__ void k_select_prefilter(const int64_t *spans, int64_t m, const int64_t *lits, int64_t l, uint8_t *sel) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  int64_t pos = 0, i = 0, j = 0;
  while (true) {
    while (j < l && lits[2 * j] < pos) ++j;
    if (j == l) break;
    const int64_t c = lits[2 * j];
    while (i < m && spans[2 * i] < c) ++i;
    if (i == m) break;
    sel[i] = 1;
    const int64_t s = spans[2 * i], e = spans[2 * i + 1];
    pos = e > s ? e : s + 1;
    ++i;
  }
}

__global__ void k_select_compact(const int64_t *spans, const uint8_t *sel, int64_t m, unsigned long long *desc,
                                 unsigned int *ticket, int64_t *out, unsigned long long *total, int n_tiles) {
  __shared__ int s_tile;
  __shared__ int ws[kThreads / 32 + 2];
  __shared__ unsigned long long s_base;
  if (threadIdx.x == 0) s_tile = (int)atomicAdd(ticket, 1u);
  __syncthreads();
  const int tile = s_tile;
  const int64_t i = (int64_t)tile * kThreads + threadIdx.x;
  const int v = (i < m && sel[i]) ? 1 : 0;
  int tot;
  const int rank = block_exclusive_scan(v, ws, &tot);
  if (threadIdx.x < 32) {
    const unsigned long long b = lookback_exclusive(desc, tile, (unsigned long long)tot);
    if (threadIdx.x == 0) {
      s_base = b;
      if (tile == n_tiles - 1) *total = b + (unsigned long long)tot;
    }
  }
  __syncthreads();
  if (v) {
    const unsigned long long o = s_base + rank;
    st_span(out + 2 * o, spans[2 * i], spans[2 * i + 1]);
  }
}

// ---------------------------------------------------------------------------------------------
// sub / split assembly.
// Segment j (0 <= j <= M): gap [pos_j, s_j) copied to out[O_j ...], then (j < M) the replacement
// (template of literal pieces and fixed-offset group slices), then for a zero-width match the
// byte text[pos_j] (matcher.mojo:1771-1778).  pos_0 = 0, pos_{j+1} = e_j (or e_j + 1 after a
// zero-width match); the tail gap j = M runs to n.
__device__ __forceinline__ int64_t seg_pos(const int64_t *spans, int64_t j, int mode) {
  if (j == 0) return 0;
  const int64_t s = spans[2 * (j - 1)], e = spans[2 * (j - 1) + 1];
  if (mode == SEG_SPLIT) return e;
  return e == s ? e + 1 : e;
}

// Output offset of every segment = exclusive scan of the segment lengths.  A CTA owns kSegTile = 2048
// consecutive segments: warp w takes the 256 from w * 256 on, in 8 rounds of 32 (lane = segment, so the
// span loads and the offset stores are coalesced); a shuffle scan per round with a running carry, a scan of
// the 8 warp totals, one decoupled look-back per CTA.
__global__ void __launch_bounds__(kThreads) k_seg_lengths(const int64_t *spans, int64_t M, int64_t n, int mode,
                                                          int64_t repl_len, unsigned long long *desc,
                                                          unsigned int *ticket, int64_t *offs /* M+2 */, int *bad,
                                                          int n_tiles, int64_t *tile_seg) {
  __shared__ int s_tile;
  __shared__ long long s_warp[kThreads / 32];
  __shared__ unsigned long long s_base;
  if (threadIdx.x == 0) s_tile = (int)atomicAdd(ticket, 1u);
  __syncthreads();
  const int tile = s_tile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t j0 = (int64_t)tile * kSegTile + (int64_t)warp * (32 * kSegRounds) + lane;  // 0..M (M = tail gap)
  long long len[kSegRounds], excl[kSegRounds];
  long long carry = 0;  // sum of this warp's earlier rounds
#pragma unroll
  for (int r = 0; r < kSegRounds; ++r) {
    const int64_t j = j0 + 32 * r;
    long long l = 0;
    if (j <= M) {
      int64_t pos = 0;
      if (j > 0) {
        const longlong2 pv = __ldg(reinterpret_cast<const longlong2 *>(spans) + (j - 1));
        pos = (mode == SEG_SPLIT) ? pv.y : (pv.y == pv.x ? pv.y + 1 : pv.y);
      }
      longlong2 cur = make_longlong2(n, n);
      if (j < M) cur = __ldg(reinterpret_cast<const longlong2 *>(spans) + j);
      long long gap = cur.x - pos;
      if (j == M && gap < 0) gap = 0;  // pos may be n + 1 after a zero-width match at n
      if (gap < 0) { if (mode == SEG_SPLIT) *bad = 1; gap = 0; }
      l = gap;
      if (j < M && mode == SEG_SUB) {
        l += repl_len;
        if (cur.x == cur.y && pos < n) l += 1;
      }
    }
    long long inc = l;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    len[r] = l;
    excl[r] = carry + inc - l;
    carry += __shfl_sync(0xffffffffu, inc, 31);
  }
  if (lane == 0) s_warp[warp] = carry;
  __syncthreads();
  long long before = 0, tot = 0;
  for (int w = 0; w < kThreads / 32; ++w) { if (w < warp) before += s_warp[w]; tot += s_warp[w]; }
  if (threadIdx.x < 32) {
    const unsigned long long b = lookback_exclusive(desc, tile, (unsigned long long)tot);
    if (threadIdx.x == 0) {
      s_base = b;
      if (tile == n_tiles - 1) offs[M + 1] = (int64_t)(b + (unsigned long long)tot);
    }
  }
  __syncthreads();
  const long long base = (long long)s_base + before;
#pragma unroll
  for (int r = 0; r < kSegRounds; ++r) {
    const int64_t j = j0 + 32 * r;
    if (j <= M) {
      const int64_t O = base + excl[r];
      offs[j] = O;
      // segment that owns the first byte of every output tile it covers (k_assemble's window)
      if (tile_seg && len[r] > 0)
        for (int64_t t = (O + kAsmTile - 1) / kAsmTile; t * kAsmTile < O + len[r]; ++t) tile_seg[t] = j;
    }
  }
}

// Output-stationary assembly, one CTA per kAsmTile (8 KiB) of output.
//
// The segments that touch the tile (tile_seg gives the window) are read once, round-robin (coalesced), into
// records {output offset, search position, gap length} in shared memory; a bitmap over the tile's kAsmTile
// output positions gets a bit where each segment WITH output bytes starts, prefix popcounts of the bitmap
// words are taken, and nzk[rank] maps the r-th set bit back to its segment.  Then every lane produces 4 consecutive output bytes per step (a warp writes 128 contiguous
// bytes, STG.32): the record of an output byte is rank(bitmap, position) -- one popcount for the first byte
// of the word, the next bit for each following byte -- so there is no search and no per-lane walk, and the
// lanes of a warp read consecutive text bytes.  Gap bytes come from the text, replacement bytes from the
// flattened template (tpl_src[r] >= 0: text[match_start + tpl_src[r]], < 0: the literal byte -1 - tpl_src[r]),
// then the extra byte of a zero-width match.  Tiles whose window exceeds the record capacity, or whose
// relative coordinates do not fit 32 bits, take the generic per-thread walk (assemble_tile_slow).
struct SegCursor {
  int64_t j, O, O_next, pos, s, gap;
};
struct AsmShared {
  uint4 rec[kAsmSegCap];          // segment lo_seg + k: x = output offset (relative to O_base), y = search
                                  //    position (relative to x_base), z = gap length
  uint16_t nzk[kAsmSegCap];       // k of the r-th segment that has output bytes (in output order)
  uint32_t bits[kAsmTile / 32];   // a set bit = such a segment starts at this output position of the tile
  uint32_t wprefix[kAsmTile / 32];
  int32_t tpl[kAsmTplCap];        // flattened replacement template
  int scan_ws[kThreads / 32 + 2];
};
__device__ __forceinline__ int64_t seg_pos_ldg(const AsmParams &p, int64_t j) {
  if (j == 0) return 0;
  const int64_t ps = __ldg(p.spans + 2 * (j - 1)), pe = __ldg(p.spans + 2 * (j - 1) + 1);
  return (p.mode == SEG_SPLIT) ? pe : (pe == ps ? pe + 1 : pe);
}
__device__ __forceinline__ void seg_load(const AsmParams &p, int64_t j, SegCursor &c) {
  c.j = j;
  c.O = __ldg(p.offs + j);
  c.O_next = __ldg(p.offs + j + 1);
  c.pos = seg_pos_ldg(p, j);
  c.s = j < p.M ? __ldg(p.spans + 2 * j) : p.n;
  c.gap = c.s > c.pos ? c.s - c.pos : 0;
}

// generic path: thread = 16 consecutive output bytes, binary search over offs, byte-by-byte walk
__device__ __forceinline__ void assemble_tile_slow(const AsmParams &p, int64_t tile, int64_t lo_seg, int64_t hi_seg) {
  uint32_t cw = 0;         // cached aligned text word (aligned in the address space: the text itself
  uintptr_t cw_addr = 1;   // may start at any byte of a wrapped tensor); a word that holds a valid byte
                           // never leaves that byte's 32-byte sector, so the over-read is harmless
  auto rd = [&](int64_t x) -> uint32_t {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p.text + x);
    const uintptr_t wa = a & ~(uintptr_t)3;
    if (wa != cw_addr) {
      cw = __ldg(reinterpret_cast<const uint32_t *>(wa));
      cw_addr = wa;
    }
    return (cw >> (8 * (int)(a & 3))) & 0xffu;
  };
#pragma unroll 1
  for (int r = 0; r < kAsmRounds; ++r) {
    const int64_t o0 = tile * kAsmTile + ((int64_t)r * kThreads + threadIdx.x) * 16;
    if (o0 >= p.total) return;
    int64_t lo = lo_seg, hi = hi_seg;  // largest j in [lo_seg, hi_seg] with offs[j] <= o0
    while (lo < hi) {
      const int64_t mid = (lo + hi + 1) >> 1;
      if (__ldg(p.offs + mid) <= o0) lo = mid; else hi = mid - 1;
    }
    SegCursor c;
    seg_load(p, lo, c);
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int64_t o = o0 + k;
      uint32_t b = 0;
      if (o < p.total) {
        while (o >= c.O_next) seg_load(p, c.j + 1, c);
        const int64_t d = o - c.O;
        if (d < c.gap) {
          b = rd(c.pos + d);
        } else {
          const int64_t q = d - c.gap;
          if (q < p.repl_len) {
            const int32_t src = __ldg(p.tpl_src + q);
            b = src >= 0 ? rd(c.s + src) : (uint32_t)(-1 - src);
          } else {
            b = rd(c.pos);  // zero-width match: the byte at the search position (matcher.mojo:1771-1778)
          }
        }
      }
      w[k >> 2] |= b << (8 * (k & 3));
    }
    *reinterpret_cast<uint4 *>(p.out + o0) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

__global__ void __launch_bounds__(kThreads) k_assemble(const AsmParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  AsmShared &sh = *reinterpret_cast<AsmShared *>(smem_raw);
  const int tid = threadIdx.x;
  const int64_t tile = blockIdx.x;
  const int64_t lo_seg = __ldg(p.tile_seg + tile);
  const int64_t hi_seg = (tile + 1 < p.n_tiles) ? __ldg(p.tile_seg + tile + 1) : p.M;
  const int64_t cnt = hi_seg - lo_seg + 1;
  const int64_t O_base = __ldg(p.offs + lo_seg), O_end = __ldg(p.offs + hi_seg + 1);
  const int64_t x_base = seg_pos_ldg(p, lo_seg);
  const int64_t x_last = hi_seg < p.M ? __ldg(p.spans + 2 * hi_seg + 1) + 1 : p.n + 1;  // bounds every pos / s of the window
  const bool fast = cnt <= kAsmSegCap && p.repl_len <= kAsmTplCap && O_end - O_base < 0xfffffff0ll &&
                    x_last - x_base < 0xfffffff0ll;
  if (!fast) {
    assemble_tile_slow(p, tile, lo_seg, hi_seg);
    return;
  }
  const int64_t tile_lo = tile * kAsmTile;
  const uint32_t tile_rel = (uint32_t)(tile_lo - O_base);  // tile start in record coordinates (>= 0: lo_seg owns it)
  const int n = (int)cnt;
  {  // the tile's text is first touched in the output loop: start pulling it into L2 now (one line per thread)
    const int64_t x = x_base + tile_rel + (int64_t)tid * 128;  // the tile starts about tile_rel bytes into lo_seg's gap
    if (x < x_last && x < p.n) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.text + x));
  }
  for (int i = tid; i < kAsmTile / 32; i += kThreads) sh.bits[i] = 0;
  for (int i = tid; i < (int)p.repl_len; i += kThreads) sh.tpl[i] = __ldg(p.tpl_src + i);
  // ---- records, one per segment of the window, loaded round-robin (coalesced)
  for (int k = tid; k < n; k += kThreads) {
    const int64_t j = lo_seg + k;
    const int64_t pj = seg_pos_ldg(p, j);
    const int64_t sj = j < p.M ? __ldg(p.spans + 2 * j) : p.n;
    sh.rec[k] = make_uint4((uint32_t)(__ldg(p.offs + j) - O_base), (uint32_t)(pj - x_base),
                           (uint32_t)(sj > pj ? sj - pj : 0), 0u);
  }
  __syncthreads();
  // ---- a bit where every segment with output bytes starts (the owner of the tile's first byte: bit 0)
  const uint32_t end_rel = (uint32_t)(O_end - O_base);
  for (int k = tid; k < n; k += kThreads) {
    const uint32_t Ok = sh.rec[k].x, On = (k + 1 < n) ? sh.rec[k + 1].x : end_rel;
    if (On > Ok && Ok < tile_rel + (uint32_t)kAsmTile) {
      const uint32_t bpos = Ok > tile_rel ? Ok - tile_rel : 0u;
      atomicOr(&sh.bits[bpos >> 5], 1u << (bpos & 31u));
    }
  }
  __syncthreads();
  // ---- exclusive prefix popcount of the bitmap words (kAsmTile / 32 words, kWpt consecutive ones per thread)
  {
    constexpr int kWpt = kAsmTile / 32 / kThreads;
    static_assert(kWpt >= 1 && kWpt * kThreads * 32 == kAsmTile, "tile must be a multiple of 32 bits per thread");
    int c[kWpt], sum = 0;
#pragma unroll
    for (int i = 0; i < kWpt; ++i) { c[i] = __popc(sh.bits[kWpt * tid + i]); sum += c[i]; }
    int tot;
    int ex = block_exclusive_scan(sum, sh.scan_ws, &tot);
#pragma unroll
    for (int i = 0; i < kWpt; ++i) { sh.wprefix[kWpt * tid + i] = (uint32_t)ex; ex += c[i]; }
  }
  __syncthreads();
  // ---- rank of every marked segment -> its k
  for (int k = tid; k < n; k += kThreads) {
    const uint32_t Ok = sh.rec[k].x, On = (k + 1 < n) ? sh.rec[k + 1].x : end_rel;
    if (On > Ok && Ok < tile_rel + (uint32_t)kAsmTile) {
      const uint32_t bpos = Ok > tile_rel ? Ok - tile_rel : 0u;
      const int r = (int)sh.wprefix[bpos >> 5] + __popc(sh.bits[bpos >> 5] & ((1u << (bpos & 31u)) - 1u));
      sh.nzk[r] = (uint16_t)k;
    }
  }
  __syncthreads();
  // ---- output: lane = 4 consecutive bytes per step, straight-line code, one text load per byte
  const uint8_t *xb = p.text + x_base;
  const uint32_t repl = (uint32_t)p.repl_len;
  const uint32_t total_rel = (uint32_t)min((int64_t)kAsmTile, p.total - tile_lo);  // live bytes of this tile
#pragma unroll 1
  for (int st = 0; st < kAsmTile / (4 * kThreads); ++st) {
    const uint32_t t0 = (uint32_t)(st * kThreads + tid) * 4u;  // tile-relative position of this lane's word
    if (t0 >= total_rel) break;
    const uint32_t bw = sh.bits[t0 >> 5];
    const int sh5 = (int)(t0 & 31u);
    int rk = (int)sh.wprefix[t0 >> 5] + __popc(bw & ((2u << sh5) - 1u)) - 1;  // record of byte 0 (>= 0)
    uint4 rc = sh.rec[sh.nzk[rk]];
    uint32_t off[4], lit[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      if (b > 0 && ((bw >> (sh5 + b)) & 1u)) rc = sh.rec[sh.nzk[++rk]];
      const uint32_t d = tile_rel + t0 + b - rc.x;  // offset inside the segment
      uint32_t o = rc.y + d;                        // gap byte
      uint32_t l = 0x100u;                          // >= 0x100: take the loaded text byte
      if (d >= rc.z) {
        const uint32_t q = d - rc.z;
        o = rc.y;                                   // zero-width match: the byte at the search position
        if (q < repl) {
          const int32_t src = sh.tpl[q];
          if (src >= 0) o = rc.y + rc.z + (uint32_t)src; else l = (uint32_t)(-1 - src);
        }
      }
      if (t0 + b >= total_rel) { o = rc.y; l = 0; }  // past the end of the output: zero
      off[b] = o;
      lit[b] = l;
    }
    uint32_t word = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const uint32_t by = lit[b] < 0x100u ? lit[b] : (uint32_t)__ldg(xb + off[b]);
      word |= by << (8 * b);
    }
    *reinterpret_cast<uint32_t *>(p.out + tile_lo + t0) = word;
  }
}

// ---------------------------------------------------------------------------------------------
// host-side launchers
template <int FILT>
static void launch_scan_t(const ScanParams &p, int grid, cudaStream_t st) {
  static bool attr_set_dev[kLaunchDevSlots] = {};
  bool &attr_set = attr_set_dev[launch_dev()];
  if (!attr_set) {
    cudaFuncSetAttribute(k_scan<FILT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScanShared));
    attr_set = true;
  }
  k_scan<FILT><<<grid, kThreads, sizeof(ScanShared), st>>>(p);
}
void launch_scan(const ScanParams &p, int grid, cudaStream_t st) {
  switch (p.filt.kind) {
    case FILT_NEEDLES: launch_scan_t<FILT_NEEDLES>(p, grid, st); break;
    case FILT_RANGE: launch_scan_t<FILT_RANGE>(p, grid, st); break;
    case FILT_ALL: launch_scan_t<FILT_ALL>(p, grid, st); break;
    default: launch_scan_t<FILT_LUT>(p, grid, st); break;
  }
}
void launch_lines(LinesParams p, cudaStream_t st) {
  if (p.n_docs <= 0) return;
  // staging: fixed-size documents, 16-byte aligned, all of them complete, rows fit next to the tables
  int pitch = (int)p.doc_stride;
  if (((pitch >> 4) & 1) == 0) pitch += 16;  // odd number of 16-byte units per row
  const int tables = p.rev_bytes + p.fwd_bytes;
  const bool staged = p.doc_offsets == nullptr && p.doc_stride >= 16 && (p.doc_stride & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(p.text) & 15) == 0 && p.n_docs * p.doc_stride <= p.n &&
                      tables + kThreads * pitch <= 56 * 1024;
  p.staged = staged ? 1 : 0;
  p.line_pitch = pitch;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  {  // candidates-only walk when the automaton has a (rare) required byte; MRX_LINES_FILTER=0 turns it off
    const char *env = getenv("MRX_LINES_FILTER");
    const size_t req_smem = sizeof(LinesShared) + (size_t)tables +
                            (size_t)kReqWarps * ((size_t)kReqRing * pitch + kReqRing * sizeof(int64_t));
    if (staged && p.req_byte >= 0 && p.doc_stride <= 16 * kReqMaxVecs && req_smem <= 220 * 1024 &&
        !(env && env[0] == '0')) {
      static size_t req_attr_dev[kLaunchDevSlots] = {};
      size_t &req_attr = req_attr_dev[launch_dev()];
      if (req_smem > req_attr) {
        cudaFuncSetAttribute(k_lines_req<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)req_smem);
        cudaFuncSetAttribute(k_lines_req<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)req_smem);
        req_attr = req_smem;
      }
      const int64_t n_groups = (p.n_docs + 31) / 32;
      const int64_t want = (n_groups + kReqWarps - 1) / kReqWarps;
      const unsigned grid = (unsigned)(want < sms ? want : sms);
      if (p.doc_stride == 128) k_lines_req<8><<<grid, kReqThreads, req_smem, st>>>(p);
      else k_lines_req<0><<<grid, kReqThreads, req_smem, st>>>(p);
      return;
    }
  }
  const int smem = (int)sizeof(LinesShared) + tables + (staged ? kThreads * pitch : 0);
  static int attr_smem_dev[kLaunchDevSlots] = {};
  int &attr_smem = attr_smem_dev[launch_dev()];
  if (smem > attr_smem) {
    cudaFuncSetAttribute(k_lines, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr_smem = smem;
  }
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_lines, kThreads, smem);
  const int64_t n_batches = (p.n_docs + kThreads - 1) / kThreads;
  const int64_t resident = (int64_t)sms * (per_sm > 0 ? per_sm : 1);
  k_lines<<<(unsigned)(n_batches < resident ? n_batches : resident), kThreads, smem, st>>>(p);
}
template <int FILT>
static void launch_runs_t(const RunsParams &p, int grid, cudaStream_t st) {
  static bool attr_set_dev[kLaunchDevSlots] = {};
  bool &attr_set = attr_set_dev[launch_dev()];
  if (!attr_set) {
    cudaFuncSetAttribute(k_runs<FILT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RunsShared));
    attr_set = true;
  }
  k_runs<FILT><<<grid, kThreads, sizeof(RunsShared), st>>>(p);
}
void launch_runs(const RunsParams &p, int grid, cudaStream_t st) {
  switch (p.filt.kind) {
    case FILT_NEEDLES: launch_runs_t<FILT_NEEDLES>(p, grid, st); break;
    case FILT_RANGE: launch_runs_t<FILT_RANGE>(p, grid, st); break;
    case FILT_ALL: launch_runs_t<FILT_ALL>(p, grid, st); break;
    default: launch_runs_t<FILT_LUT>(p, grid, st); break;
  }
}
void launch_dense(const DenseParams &p, int grid, cudaStream_t st) {
  const int smem = (int)sizeof(DenseShared) + p.table_bytes;
  static int attr_smem_dev[kLaunchDevSlots] = {}, resident_dev[kLaunchDevSlots] = {};
  int &attr_smem = attr_smem_dev[launch_dev()], &resident = resident_dev[launch_dev()];
  if (smem > attr_smem) {
    cudaFuncSetAttribute(k_dense, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_dense_first, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr_smem = smem;
    int per_sm = 0, dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_dense, kThreads, smem);
    resident = sms * (per_sm > 0 ? per_sm : 1);
  }
  if (p.out_kind == OUT_FIRST) {
    k_dense_first<<<grid, kThreads, smem, st>>>(p);
  } else {
    // persistent: every CTA must be resident (tiles are handed out by ticket, flushes are deferred)
    k_dense<<<grid < resident ? grid : resident, kThreads, smem, st>>>(p);
  }
}
void launch_tile_trailing(const uint8_t *text, int64_t n, int64_t origin, int n_tiles, const DevProg *prog,
                          long long *trail, int *all, cudaStream_t st) {
  const int warps_per_block = 8;
  k_tile_trailing<<<(n_tiles + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, st>>>(
      text, n, origin, n_tiles, prog, trail, all);
}
void launch_carry_scan(int n_tiles, const long long *trail, const int *all, long long *carry, const long long *carry0,
                       cudaStream_t st) {
  k_carry_scan<<<1, 1024, 0, st>>>(n_tiles, trail, all, carry, carry0);
}
void launch_back_run(const uint8_t *text, int64_t origin, const DevProg *prog, long long *out, cudaStream_t st) {
  k_back_run<<<1, kThreads, 0, st>>>(text, origin, prog, out);
}
void launch_batch_trivial(const ScanParams &p, int wildcard, int64_t *out, cudaStream_t st) {
  if (p.n_docs <= 0) return;
  k_batch_trivial<<<(unsigned)((p.n_docs + 255) / 256), 256, 0, st>>>(p, wildcard, out);
}
void launch_attempt_one(const ScanParams &p, int64_t pos, int64_t *out, cudaStream_t st) {
  k_attempt_one<<<1, 32, 0, st>>>(p, pos, out);
}
void launch_batch_finalize(const unsigned long long *doc_best, int64_t n_docs, const int64_t *doc_offsets,
                           int64_t doc_stride, int64_t n, int empty_at_end, int64_t *out, cudaStream_t st) {
  k_batch_finalize<<<(unsigned)((n_docs + 255) / 256), 256, 0, st>>>(doc_best, n_docs, doc_offsets, doc_stride, n,
                                                                     empty_at_end, out);
}
void launch_add_i64(int64_t *p, int64_t n, int64_t v, cudaStream_t st) {
  if (n <= 0) return;
  k_add_i64<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, n, v);
}
void launch_fill_u64(unsigned long long *p, int64_t n, unsigned long long v, cudaStream_t st) {
  k_fill_u64<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, n, v);
}
void launch_select_prefix_max(const int64_t *spans, int64_t m, unsigned long long *desc, unsigned int *ticket,
                              int64_t *pmax, cudaStream_t st) {
  k_select_prefix_max<<<(unsigned)((m + kThreads - 1) / kThreads), kThreads, 0, st>>>(spans, m, desc, ticket, pmax);
}
void launch_select_walk(const int64_t *spans, const int64_t *keys, const int64_t *pmax, int64_t m, uint8_t *sel,
                        cudaStream_t st) {
  k_select_walk<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(spans, keys, pmax, m, sel);
}
void launch_select_prefilter(const int64_t *spans, int64_t m, const int64_t *lits, int64_t l, uint8_t *sel, cudaStream_t st) {
  k_select_prefilter<<<1, 32, 0, st>>>(spans, m, lits, l, sel);
}
void launch_select_compact(const int64_t *spans, const uint8_t *sel, int64_t m, unsigned long long *desc,
                           unsigned int *ticket, int64_t *out, unsigned long long *total, cudaStream_t st) {
  const int tiles = (int)((m + kThreads - 1) / kThreads);
  k_select_compact<<<tiles, kThreads, 0, st>>>(spans, sel, m, desc, ticket, out, total, tiles);
}
void launch_seg_lengths(const int64_t *spans, int64_t M, int64_t n, int mode, int64_t repl_len,
                        unsigned long long *desc, unsigned int *ticket, int64_t *offs, int *bad, int64_t *tile_seg,
                        cudaStream_t st) {
  const int tiles = (int)((M + 1 + kSegTile - 1) / kSegTile);
  k_seg_lengths<<<tiles, kThreads, 0, st>>>(spans, M, n, mode, repl_len, desc, ticket, offs, bad, tiles, tile_seg);
}
void launch_assemble(const AsmParams &p, cudaStream_t st) {
  if (p.total <= 0) return;
  static bool attr_set_dev[kLaunchDevSlots] = {};
  bool &attr_set = attr_set_dev[launch_dev()];
  if (!attr_set) {
    cudaFuncSetAttribute(k_assemble, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AsmShared));
    attr_set = true;
  }
  k_assemble<<<(unsigned)p.n_tiles, kThreads, sizeof(AsmShared), st>>>(p);
}

}  // namespace mrx
