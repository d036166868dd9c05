// k_subst<FILT> -- sub() for chain automata over one class with a fixed width (BASELINE config 4:
// `(\d{3})(\d{3})(\d{4})` -> `\1-\2-\3`) as ONE pass over the text: matches are found, the replacement template is
// expanded and the rewritten text is written without a match list or an offset array ever reaching HBM.  (The general
// path runs k_runs -> 16 B per match -> k_seg_lengths -> k_assemble, whose per-output-byte source selection costs
// 1.7 G warp-instructions per GiB: 13 ms per 4 GiB, 0.105 of the HBM roofline.)
//
// Reference semantics: _sub_impl_with_repl, fixed-width branch (matcher.mojo:1696-1781) over match_next restarts
// (dfa.mojo:1874-1903): inside a maximal class run that starts at q the matches are [q + W k, q + W k + W) while W
// more bytes of the run remain.
//
// Every match has the same width W and every replacement the same length R, so the output position of a text byte
// is p + (R - W) * (matches that start before p): one counter to scan.  A CTA owns a 36 864-byte tile, thread t its
// bytes [144 t, 144 t + 144):
//   1. the tile with 64 bytes after it is staged by one TMA bulk copy;
//   2. class bits of the chunk (SWAR range / needle test or a 256-byte table); the run length entering the chunk by a
//      segmented scan over the block seeded with the tile's carry (k_tile_trailing + k_carry_scan, as for findall);
//      the bits of the 32 bytes after the chunk from the neighbour;
//   3. match starts, bit-parallel: E = positions followed by W in-class bytes (AND-smear), seeds = run starts in E
//      (for a run that continues from before the chunk: its first start at the run's phase), then
//      M |= (M << W) & E until nothing changes -- a hop stays inside the run because both ends see W class bytes;
//   4. match counts are scanned across the block and published as the tile's aggregate; the descriptor loads of the
//      look-back window go out here and are consumed after step 5;
//   5. every thread writes the bytes of its chunk that lie outside matches into a staging buffer in shared memory at
//      (position + (R - W) * starts before) and its starts into the tile's match list; then the block expands the
//      template match by match (literal bytes, or bytes of the match at the group's offset read from the staged text);
//   6. with the tile's base known the staging buffer is copied to its place in the output with 16-byte stores while
//      the next tile's text is already in flight.
// Tiles by ticket; persistent CTAs, every CTA resident (a tile waits for the aggregates of the tiles before it).
#include <cstddef>
#include <cstdint>

#include "kernels.cuh"

namespace mrx {

constexpr int32_t kTplSkip = INT32_MIN;  // template entry past the replacement's end
constexpr int kSubstPost = 64;  // staged bytes after the tile (W <= 32: a match reaches at most 31 bytes over)

struct SubstShared {
  alignas(128) uint8_t text[kTile + kSubstPost];
  uint8_t lut[256];            // class membership (FILT_LUT only)
  alignas(16) int32_t tpl[kSubstMaxRepl];  // flattened template: >= 0: byte of the match at this offset, < 0: literal -1 - value
  uint32_t head_bits[kThreads + 1];
  alignas(8) uint64_t bar;
  unsigned long long red_ws[kThreads / 32];
  long long warp_len[kThreads / 32];
  int warp_all[kThreads / 32];
  int scan_ws[kThreads / 32];
  int next_tile, skip0, over;
  alignas(16) uint8_t stage[16];  // the rewritten tile (SubstParams::stage_bytes of dynamic shared memory)
};

template <int FILT>
__device__ __forceinline__ uint32_t subst_bits16(const uint4 &w, const Filter &f, const uint8_t *lut) {
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
    } else if (FILT == FILT_RANGE) {
      fl = swar_range(ws[i], f.n0, f.n1, f.hi_is_ff != 0);
    } else {
      fl = (lut[ws[i] & 255] ? 0x80u : 0) | (lut[(ws[i] >> 8) & 255] ? 0x8000u : 0) | (lut[(ws[i] >> 16) & 255] ? 0x800000u : 0) |
           (lut[ws[i] >> 24] ? 0x80000000u : 0);
    }
    m |= flags_to_nibble(fl) << (4 * i);
  }
  return m;
}

__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// out[16 i, +16) = bytes [4 HW + sh8 / 8 + 16 i, +16) of the staging buffer: two conflict-free 16-byte loads each
template <int HW>
__device__ __forceinline__ void subst_copy_body(const uint4 *sv, uint8_t *dst, int body, int sh8, int tid) {
  for (int i = tid; i < body; i += kThreads) {
    const uint4 a = sv[i], b = sv[i + 1];
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint4 o;
    o.x = __funnelshift_r(w[HW], w[HW + 1], sh8); o.y = __funnelshift_r(w[HW + 1], w[HW + 2], sh8);
    o.z = __funnelshift_r(w[HW + 2], w[HW + 3], sh8); o.w = __funnelshift_r(w[HW + 3], w[HW + 4], sh8);
    *reinterpret_cast<uint4 *>(dst + 16 * i) = o;
  }
}

template <int FILT>
__global__ void __launch_bounds__(kThreads, 2) k_subst(const SubstParams prm) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  SubstShared &sh = *reinterpret_cast<SubstShared *>(smem_raw);
  constexpr int W8 = kThreads / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool text_aligned = (reinterpret_cast<uintptr_t>(prm.text) & 15) == 0;
  const int W = prm.width, R = prm.repl_len, delta = R - W;
  uint16_t *list = reinterpret_cast<uint16_t *>(sh.stage + prm.stage_bytes);  // tile-relative starts of the tile's matches
  if (tid == 0) {
    mbar_init(&sh.bar, 1);
    fence_proxy_async();
  }
  sh.lut[tid] = prm.prog->cand[tid];  // kThreads == 256
  if (tid < kSubstMaxRepl) sh.tpl[tid] = tid < R ? prm.tpl[tid] : kTplSkip;
  __syncthreads();

  auto issue_load = [&](int t) {  // thread 0 only
    if (t >= prm.n_tiles) return;
    const int64_t ts = (int64_t)t * kTile;
    int64_t bytes = min((int64_t)kTile + kSubstPost, prm.n - ts);
    if (bytes < 0) bytes = 0;
    const uint32_t bulk = text_aligned ? ((uint32_t)bytes & ~15u) : 0u;
    mbar_expect_tx(&sh.bar, bulk);
    if (bulk) tma_bulk_g2s(sh.text, prm.text + ts, bulk, &sh.bar);
  };

  unsigned long long base_m = 0;  // matches that start before tile `covered`
  int covered = 0;
  unsigned long long win[2];
  auto window_load = [&](int p) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int i = p - 1 - (tid + kThreads * u);
      win[u] = i >= covered ? ld_volatile_u64(prm.desc + i) : kFlagAgg;
    }
  };
  auto window_sum = [&](int p) -> unsigned long long {  // this warp's part
    unsigned long long s = 0;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int i = p - 1 - (tid + kThreads * u);
      while ((win[u] >> 62) == 0) win[u] = ld_volatile_u64(prm.desc + i);
      s += win[u] & kValMask;
    }
    for (int i = p - 1 - 2 * kThreads - tid; i >= covered; i -= kThreads) {  // (rare) the rest of a long window
      unsigned long long d;
      do { d = ld_volatile_u64(prm.desc + i); } while ((d >> 62) == 0);
      s += d & kValMask;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
  };

  if (tid == 0) {
    sh.next_tile = (int)atomicAdd(prm.ticket, 1u);
    issue_load(sh.next_tile);
  }
  __syncthreads();
  int tile = sh.next_tile;
  for (int it = 0; tile < prm.n_tiles; ++it) {
    int ticket = 0;
    if (tid == 0) ticket = (int)atomicAdd(prm.ticket, 1u);  // the next tile; posted before barrier #3
    const int64_t tile_start = (int64_t)tile * kTile;
    const int64_t left = prm.n - tile_start;
    const int valid = (int)min((int64_t)kTile, left);
    const int staged = (int)min((int64_t)kTile + kSubstPost, left);
    {
      const int bulk = text_aligned ? (staged & ~15) : 0;
      for (int i = bulk + tid; i < staged; i += kThreads) sh.text[i] = prm.text[tile_start + i];
      mbar_wait(&sh.bar, (uint32_t)(it & 1));
      if (bulk != staged) __syncthreads();
    }
    // ---- 2. class bits of my chunk (bytes past the text: not in class)
    const int chunk_rel = tid * kChunk;
    const int cv = max(0, min(kChunk, valid - chunk_rel));
    uint32_t cb[5] = {0, 0, 0, 0, 0};
    uint4 mine[kVecs];  // my chunk's text, kept for step 5
#pragma unroll
    for (int v = 0; v < kVecs; ++v) {
      const uint4 tv = mine[v] = *reinterpret_cast<const uint4 *>(sh.text + chunk_rel + 16 * v);
      uint32_t m = subst_bits16<FILT>(tv, prm.filt, sh.lut);
      const int live = cv - 16 * v;
      if (live < 16) m &= live <= 0 ? 0u : ((1u << live) - 1u);
      cb[v >> 1] |= m << (16 * (v & 1));
    }
    int lead = 0;  // leading / trailing in-class bytes of the chunk
    {
      bool all = true;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const uint32_t full = i == 4 ? 0xffffu : 0xffffffffu;
        if (all) {
          if (cb[i] == full) lead += i == 4 ? 16 : 32;
          else { lead += __ffs((int)~cb[i]) - 1; all = false; }
        }
      }
    }
    const bool all_in = lead == kChunk;
    int trail = 0;
    if (!all_in) {
      bool go = true;
#pragma unroll
      for (int i = 4; i >= 0; --i) {
        const int bits = i == 4 ? 16 : 32;
        const uint32_t full = i == 4 ? 0xffffu : 0xffffffffu;
        if (go) {
          if (cb[i] == full) trail += bits;
          else { trail += __clz((int)~(cb[i] << (32 - bits))); go = false; }
        }
      }
    }
    // inclusive segmented scan of (all, len): (a1, l1) * (a2, l2) = (a1 & a2, a2 ? l1 + l2 : l2)
    long long len = all_in ? kChunk : trail;
    int allf = all_in ? 1 : 0;
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
    sh.head_bits[tid] = cb[0];
    if (tid == kThreads - 1) {  // the 32 bytes after the tile, from the staged look-ahead
      uint32_t hb = 0;
      const int post = staged - kTile;
      if (post > 0) {
        const uint4 a = *reinterpret_cast<const uint4 *>(sh.text + kTile), b = *reinterpret_cast<const uint4 *>(sh.text + kTile + 16);
        hb = subst_bits16<FILT>(a, prm.filt, sh.lut) | (subst_bits16<FILT>(b, prm.filt, sh.lut) << 16);
        if (post < 32) hb &= (1u << post) - 1u;
      }
      sh.head_bits[kThreads] = hb;
    }
    __syncthreads();  // #1
    long long carry = prm.tile_carry[tile];  // in-class bytes immediately before the tile
    for (int w = 0; w < warp; ++w) carry = sh.warp_all[w] ? carry + sh.warp_len[w] : sh.warp_len[w];
    {
      const long long pl = __shfl_up_sync(0xffffffffu, len, 1);
      const int pa = __shfl_up_sync(0xffffffffu, allf, 1);
      if (lane > 0) carry = pa ? carry + pl : pl;
    }
    // carry = consecutive in-class bytes immediately before my chunk
    const uint32_t next_head = sh.head_bits[tid + 1];
    // ---- 3. match starts.  Window = my 144 bits + the 32 after them.
    uint32_t E[6] = {cb[0], cb[1], cb[2], cb[3], cb[4] | (next_head << 16), next_head >> 16};
    {
      int p = 1;
      auto and_shr = [&](int s) {
#pragma unroll
        for (int k = 0; k < 6; ++k) E[k] &= __funnelshift_r(E[k], k < 5 ? E[k + 1] : 0u, s);
      };
      while (2 * p <= W) { and_shr(p); p *= 2; }
      if (W > p) and_shr(W - p);
    }
    E[4] &= 0xffffu;
    uint32_t M[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const uint32_t before = k ? cb[k - 1] >> 31 : (carry > 0 ? 1u : 0u);
      M[k] = cb[k] & ~((cb[k] << 1) | before) & E[k];
    }
    int skip = 0;  // bytes at the start of my chunk that belong to a match begun before it
    if (carry > 0 && (cb[0] & 1u)) {
      const int ext = next_head == 0xffffffffu ? 32 : __ffs((int)~next_head) - 1;
      const int lead_total = lead + (all_in ? ext : 0);
      const int ph = carry < 0x7fffffffll ? (int)((unsigned)carry % (unsigned)W) : (int)(carry % W);
      const int r0 = ph ? W - ph : 0;
      if (ph && lead_total >= r0) skip = r0;
      if (r0 + W <= lead_total) M[0] |= 1u << r0;
    }
    for (;;) {
      uint32_t any = 0;
#pragma unroll
      for (int k = 4; k >= 0; --k) {
        const uint32_t t = __funnelshift_lc(k ? M[k - 1] : 0u, M[k], W) & E[k] & ~M[k];
        M[k] |= t;
        any |= t;
      }
      if (!__any_sync(0xffffffffu, any != 0)) break;
    }
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) cnt += __popc(M[k]);
    if (tid == 0) sh.skip0 = skip;
    if (tid == kThreads - 1) {  // bytes of the tile's last match that lie past the tile
      int o = 0;
      bool found = false;
#pragma unroll
      for (int k = 4; k >= 0; --k)
        if (!found && M[k]) { o = 32 * k + 31 - __clz((int)M[k]) + W - kChunk; found = true; }
      sh.over = o > 0 ? o : 0;
    }
    // ---- 4. ranks inside the tile, the tile's aggregate, the look-back window's loads
    int inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    if (lane == 31) sh.scan_ws[warp] = inc;
    __syncthreads();  // #2
    int rank = inc - cnt, total = 0;
#pragma unroll
    for (int w = 0; w < W8; ++w) {
      const int c = sh.scan_ws[w];
      if (w < warp) rank += c;
      total += c;
    }
    const int sk0 = sh.skip0;
    if (tid == 0) st_volatile_u64(prm.desc + tile, kFlagAgg | (unsigned long long)total);
    window_load(tile);
    // ---- 5. rewrite my chunk.  Staging offset of chunk byte i outside matches: base0 + i + delta * (starts before i)
    const int base0 = chunk_rel - sk0 + delta * rank;
    {
      uint32_t cov[5];  // bytes inside matches, and bytes past the text
#pragma unroll
      for (int k = 0; k < 5; ++k) cov[k] = M[k];
      int p = 1;
      auto or_shl = [&](int s) {
#pragma unroll
        for (int k = 4; k >= 0; --k) cov[k] |= __funnelshift_l(k ? cov[k - 1] : 0u, cov[k], s);
      };
      while (2 * p <= W) { or_shl(p); p *= 2; }
      if (W > p) or_shl(W - p);
      cov[0] |= (1u << skip) - 1u;
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const int live = cv - 32 * k;
        cov[k] |= live >= 32 ? 0u : (live <= 0 ? 0xffffffffu : ~((1u << live) - 1u));
      }
      // my starts -> the tile's match list, in order (the template pass below is spread over the block by match)
      {
        uint16_t *lp = list + rank;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          uint32_t m = M[k];
          while (m) {
            *lp++ = (uint16_t)(chunk_rel + 32 * k + __ffs((int)m) - 1);
            m &= m - 1;
          }
        }
      }
      // bytes outside matches: one byte store each.  The offset register is kept opaque so that every store is
      // [register + constant] (left to itself the compiler re-derives it as a sum of selects at every store).
      uint32_t o = smem_u32(sh.stage) + (uint32_t)base0;
#pragma unroll
      for (int v = 0; v < kVecs; ++v) {
        const uint32_t ws[4] = {mine[v].x, mine[v].y, mine[v].z, mine[v].w};
#pragma unroll
        for (int wi = 0; wi < 4; ++wi) {
          const int k = 4 * v + wi, bit = 4 * (k & 7);
          const uint32_t gone = cov[k >> 3], st = M[k >> 3];
          if (!(gone & (1u << bit))) sts_u8(o + 4 * k, ws[wi]);
          if (!(gone & (2u << bit))) sts_u8(o + 4 * k + 1, ws[wi] >> 8);
          if (!(gone & (4u << bit))) sts_u8(o + 4 * k + 2, ws[wi] >> 16);
          if (!(gone & (8u << bit))) sts_u8(o + 4 * k + 3, ws[wi] >> 24);
          if (st & (15u << bit)) o += delta;  // W >= 4: one start per word at most, the rest of the word is inside it
          asm volatile("" : "+r"(o));
        }
      }
    }
    __syncthreads();  // #2b: the match list is complete
    // ---- 5b. the templates, one match per thread at a time: stage offset = position - sk0 + delta * (index in the tile)
    for (int m = tid; m < total; m += kThreads) {
      const int s = list[m];
      const uint8_t *src0 = sh.text + s;
      uint8_t *dst = sh.stage + (s - sk0 + delta * m);
      for (int r = 0; r < R; r += 4) {  // (entries past R: kTplSkip)
        const int4 t4 = *reinterpret_cast<const int4 *>(sh.tpl + r);
        const int32_t src[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (src[q] != kTplSkip) dst[r + q] = src[q] >= 0 ? src0[src[q]] : (uint8_t)(-1 - src[q]);
      }
    }
    const unsigned long long wsum = window_sum(tile);
    if (lane == 0) sh.red_ws[warp] = wsum;
    if (tid == 0) sh.next_tile = ticket;
    __syncthreads();  // #3: the rewritten tile is complete, the text buffer is free, window sums and ticket posted
    if (tid == 0) issue_load(sh.next_tile);
    unsigned long long bm = base_m;
#pragma unroll
    for (int w = 0; w < W8; ++w) bm += sh.red_ws[w];
    base_m = bm + (unsigned long long)total;
    covered = tile + 1;
    // the tile's output: its text bytes from sk0 on and the bytes its last match takes from the next tile's text,
    // W -> R bytes per match
    const int nout = valid - sk0 + delta * total + (valid == kTile ? sh.over : 0);
    const long long out_at = tile_start + sk0 + (long long)delta * (long long)bm;
    if (tid == 0 && tile == prm.n_tiles - 1) { prm.totals[0] = (unsigned long long)(out_at + nout); prm.totals[1] = base_m; }
    // ---- 6. staging buffer -> out[out_at, out_at + nout): head up to 16-byte alignment, 16-byte body, tail
    {
      uint8_t *dst = prm.out + out_at;
      int head = (int)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15);
      if (head > nout) head = nout;
      if (tid < head) dst[tid] = sh.stage[tid];
      const int body = (nout - head) >> 4;
      // 16 output bytes = stage[head + 16 i, +16): the aligned vectors i and i + 1, shifted by `head` bytes (uniform)
      const uint4 *sv = reinterpret_cast<const uint4 *>(sh.stage);
      uint8_t *body_dst = dst + head;
      switch (head >> 2) {
        case 0: subst_copy_body<0>(sv, body_dst, body, (head & 3) * 8, tid); break;
        case 1: subst_copy_body<1>(sv, body_dst, body, (head & 3) * 8, tid); break;
        case 2: subst_copy_body<2>(sv, body_dst, body, (head & 3) * 8, tid); break;
        default: subst_copy_body<3>(sv, body_dst, body, (head & 3) * 8, tid); break;
      }
      const int done = head + 16 * body;
      if (tid < nout - done) dst[done + tid] = sh.stage[done + tid];
    }
    // no barrier here: nobody writes the staging buffer, red_ws, skip0 or next_tile again before barrier #2 of the
    // next tile, and barrier #1 of the next tile is reached by every thread only after its copy-out
    tile = sh.next_tile;
  }
}

template <int FILT>
static void launch_subst_t(const SubstParams &p, cudaStream_t st) {
  static int attr_dev[kLaunchDevSlots] = {};
  int &attr = attr_dev[launch_dev()];
  const int smem = (int)offsetof(SubstShared, stage) + p.stage_bytes + ((2 * (kTile / p.width + 2) + 15) & ~15);
  if (attr < smem) {
    cudaFuncSetAttribute(k_subst<FILT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr = smem;
  }
  int per_sm = 0, dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_subst<FILT>, kThreads, smem);
  const int resident = sms * (per_sm > 0 ? per_sm : 1);
  const int grid = p.n_tiles < resident ? p.n_tiles : resident;
  if (grid > 0) k_subst<FILT><<<grid, kThreads, smem, st>>>(p);
}

void launch_subst(const SubstParams &p, cudaStream_t st) {
  switch (p.filt.kind) {
    case FILT_NEEDLES: launch_subst_t<FILT_NEEDLES>(p, st); break;
    case FILT_RANGE: launch_subst_t<FILT_RANGE>(p, st); break;
    default: launch_subst_t<FILT_LUT>(p, st); break;
  }
}

}  // namespace mrx
