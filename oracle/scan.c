/* ORACLE (test infrastructure, NOT product code) -- scan loops.
 *
 * Plain-C restatement of the text-touching loops of msaelices/mojo-regex v0.21.0 for
 * DFA-routed patterns.  Tables come from oracle/frontend.py (256 x int32 rows, -1 = no
 * transition, exactly the reference's DFAState layout, dfa.mojo:215-254).  Every function
 * cites the reference file:line it follows (relative to /root/reference/src/regex/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  Nothing under mojo_regex_b200/ links or calls it.
 *
 * Parity status: see oracle/frontend.py header ("parity partially unpinned").
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  /* DFAEngine (dfa.mojo:257-294) */
  const int32_t *trans; /* n_states x 256 */
  const uint8_t *accepting;
  int32_t n_states, start_state;
  int32_t has_start_anchor, has_end_anchor, is_pure_literal, has_matcher, scan_eligible;
  const uint8_t *literal;
  int64_t literal_len;
  uint8_t lookup[256];    /* CharacterClassSIMD.lookup_table (simd_ops.mojo:266) */
  uint8_t simd_cand[256]; /* what the SIMD body of find_first_nibble_match accepts */
  int32_t simd_width;     /* SIMD_WIDTH of the modelled host (simd_ops.mojo:58), 16 or 32 */
  /* HybridMatcher (matcher.mojo:535-565) */
  int32_t is_wildcard_any, is_exact_literal, has_anchors_info, required_byte;
  const uint8_t *exact_literal;
  int64_t exact_len;
  const uint8_t *prefilter; /* NULL when no MemchrPrefilter */
  int64_t prefilter_len;
  /* CompiledRegex fixed-width sub metadata (matcher.mojo:946-962) */
  int32_t fixed_total_width, fixed_num_groups, fixed_concat;
  int32_t fixed_offsets[10], fixed_widths[10];
  /* AVX2 forms of lookup[] and simd_cand[] (oracle/__init__.py fills them): per LOW nibble the membership bits of
   * the high nibbles 0-7 ([0..15]) and 8-15 ([16..31]).  use_avx2 = 0 keeps the scalar loops (the tests hold the
   * two equal); the host must support AVX2 as well. */
  uint8_t nib_lookup[32], nib_simd[32];
  int32_t use_avx2;
} orx_regex;

typedef struct {
  int64_t *v; /* pairs (start,end) */
  int64_t n, cap;
} orx_spans;

static void spans_push(orx_spans *s, int64_t a, int64_t b) {
  if (s->n == s->cap) {
    s->cap = s->cap ? s->cap * 2 : 8; /* MatchList: first alloc 8, doubling (matching.mojo:140-151) */
    s->v = (int64_t *)realloc(s->v, (size_t)s->cap * 2 * sizeof(int64_t));
  }
  s->v[2 * s->n] = a;
  s->v[2 * s->n + 1] = b;
  s->n++;
}

void orx_free(void *p) { free(p); }

/* String.find (Mojo stdlib): first occurrence of needle at or after start, -1 if none. */
static int64_t str_find(const uint8_t *t, int64_t n, const uint8_t *needle, int64_t m, int64_t start) {
  if (start < 0) start = 0;
  if (m == 0) return start <= n ? start : -1;
  for (int64_t p = start; p + m <= n; ++p) {
    const uint8_t *q = (const uint8_t *)memchr(t + p, needle[0], (size_t)(n - m + 1 - p));
    if (!q) return -1;
    p = q - t;
    if (memcmp(q, needle, (size_t)m) == 0) return p;
  }
  return -1;
}

/* simd_find_byte, simd_ops.mojo:216-241 */
static int64_t simd_find_byte(const uint8_t *t, int64_t start, int64_t n, uint8_t needle) {
  if (start >= n) return -1;
  const uint8_t *q = (const uint8_t *)memchr(t + start, needle, (size_t)(n - start));
  return q ? (int64_t)(q - t) : -1;
}

/* verify_match, simd_ops.mojo:937-960 */
static int verify_match(const uint8_t *pat, int64_t m, const uint8_t *t, int64_t n, int64_t pos) {
  if (pos + m > n) return 0;
  return memcmp(t + pos, pat, (size_t)m) == 0;
}

/* simd_search, simd_ops.mojo:963-1024 (first occurrence at or after start) */
static int64_t simd_search(const uint8_t *pat, int64_t m, const uint8_t *t, int64_t n, int64_t start) {
  if (m == 0) return start;
  if (m == 1) return simd_find_byte(t, start, n, pat[0]);
  for (int64_t pos = start; pos + m <= n; ++pos)
    if (t[pos] == pat[0] && t[pos + m - 1] == pat[m - 1] && verify_match(pat, m, t, n, pos)) return pos;
  return -1;
}

#if defined(__x86_64__)
#include <immintrin.h>
/* The reference's loops are SIMD (SIMD_WIDTH bytes per step through two nibble shuffles, simd_ops.mojo:565-786); a
 * scalar port would flatter the GPU.  Exact membership of 32 bytes in an arbitrary byte set: row = nib[low nibble]
 * (bits of the high nibbles 0-7, or 8-15), tested against 1 << (high nibble & 7). */
__attribute__((target("avx2"))) static inline uint32_t member32(const uint8_t *nib, const uint8_t *p) {
  const __m256i v = _mm256_loadu_si256((const __m256i *)p);
  const __m256i a_tab = _mm256_broadcastsi128_si256(_mm_loadu_si128((const __m128i *)nib));
  const __m256i b_tab = _mm256_broadcastsi128_si256(_mm_loadu_si128((const __m128i *)(nib + 16)));
  const __m256i bit_tab = _mm256_setr_epi8(1, 2, 4, 8, 16, 32, 64, -128, 1, 2, 4, 8, 16, 32, 64, -128,
                                           1, 2, 4, 8, 16, 32, 64, -128, 1, 2, 4, 8, 16, 32, 64, -128);
  const __m256i low = _mm256_and_si256(v, _mm256_set1_epi8(0x0f));
  const __m256i high = _mm256_and_si256(_mm256_srli_epi16(v, 4), _mm256_set1_epi8(0x0f));
  const __m256i row = _mm256_blendv_epi8(_mm256_shuffle_epi8(a_tab, low), _mm256_shuffle_epi8(b_tab, low), v);
  const __m256i hit = _mm256_and_si256(row, _mm256_shuffle_epi8(bit_tab, high));
  return ~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(hit, _mm256_setzero_si256()));
}
/* first position in [pos, end) whose byte is (want = 1) / is not (want = 0) in the set; end if none */
__attribute__((target("avx2"))) static int64_t scan_set_avx2(const uint8_t *nib, const uint8_t *tab, const uint8_t *t,
                                                             int64_t pos, int64_t end, int want) {
  for (int k = 0; k < 16 && pos < end; ++k, ++pos) /* the next hit is usually a few bytes away */
    if ((tab[t[pos]] != 0) == (want != 0)) return pos;
  while (pos + 32 <= end) {
    uint32_t m = member32(nib, t + pos);
    if (!want) m = ~m;
    if (m) return pos + __builtin_ctz(m);
    pos += 32;
  }
  while (pos < end && (tab[t[pos]] != 0) != (want != 0)) ++pos;
  return pos;
}
static int have_avx2(void) {
  static int v = -1;
  if (v < 0) v = __builtin_cpu_supports("avx2") ? 1 : 0;
  return v;
}
#else
static int have_avx2(void) { return 0; }
static int64_t scan_set_avx2(const uint8_t *nib, const uint8_t *tab, const uint8_t *t, int64_t pos, int64_t end, int want) {
  (void)nib; (void)tab; (void)t; (void)end; (void)want;
  return pos;
}
#endif

/* CharacterClassSIMD.find_first_nibble_match, simd_ops.mojo:565-648.  The vector body
 * accepts simd_cand[] (exact for 1..3 ranges, nibble-product otherwise, :117-126); the
 * scalar tail -- the last (n - start) % SIMD_WIDTH bytes -- uses the exact lookup. */
static int64_t find_first_nibble_match(const orx_regex *re, const uint8_t *t, int64_t start, int64_t n) {
  int64_t pos = start;
  const int64_t w = re->simd_width;
  if (re->use_avx2 && have_avx2() && start < n) {
    const int64_t body_end = start + (n - start) / w * w;  /* where the reference's vector body stops */
    pos = scan_set_avx2(re->nib_simd, re->simd_cand, t, start, body_end, 1);
    if (pos < body_end) return pos;
    pos = scan_set_avx2(re->nib_lookup, re->lookup, t, body_end, n, 1);
    return pos < n ? pos : -1;
  }
  while (pos + w <= n) {
    for (int64_t i = 0; i < w; ++i)
      if (re->simd_cand[t[pos + i]]) return pos + i;
    pos += w;
  }
  while (pos < n) {
    if (re->lookup[t[pos]]) return pos;
    ++pos;
  }
  return -1;
}

/* CharacterClassSIMD.count_consecutive_matches, simd_ops.mojo:650-786 (exact lookup) */
static int64_t count_consecutive_matches(const orx_regex *re, const uint8_t *t, int64_t start, int64_t n) {
  int64_t pos = start;
  if (re->use_avx2 && have_avx2() && start < n) return scan_set_avx2(re->nib_lookup, re->lookup, t, start, n, 0) - start;
  while (pos < n && re->lookup[t[pos]]) ++pos;
  return pos - start;
}

/* DFAEngine._try_match_simd, dfa.mojo:2132-2197 */
static int try_match_simd(const orx_regex *re, const uint8_t *t, int64_t n, int64_t start, int64_t *e) {
  if (!re->has_matcher || re->n_states == 0) return 0;
  int start_acc = re->accepting[re->start_state];
  if (!start_acc && !re->scan_eligible) return 0;
  int64_t cnt = count_consecutive_matches(re, t, start, n);
  int64_t end = start + cnt;
  int valid = 0;
  if (cnt == 0) {
    if (start_acc) { valid = 1; end = start; }
  } else {
    valid = 1;
  }
  if (valid) {
    if (re->has_end_anchor && end != n) return 0;
    *e = end;
    return 1;
  }
  return 0;
}

/* DFAEngine._try_match_at_position, dfa.mojo:1905-2026.  Returns 1 and (s,e) on match. */
static int try_match_at_position(const orx_regex *re, const uint8_t *t, int64_t n, int64_t start,
                                 int exact, int64_t *s, int64_t *e) {
  if (start > n) return 0;
  if (re->is_pure_literal) { /* :1930-1951 */
    if (exact) {
      if (verify_match(re->literal, re->literal_len, t, n, start)) {
        *s = start; *e = start + re->literal_len; return 1;
      }
      return 0;
    }
    int64_t pos = simd_search(re->literal, re->literal_len, t, n, start);
    if (pos != -1) { *s = pos; *e = pos + re->literal_len; return 1; }
    return 0;
  }
  if (re->has_matcher && re->n_states > 0 &&
      (re->scan_eligible || re->accepting[re->start_state])) { /* :1956-1968 */
    int64_t end;
    if (try_match_simd(re, t, n, start, &end)) { *s = start; *e = end; return 1; }
  }
  if (start == n) { /* :1970-1977 */
    if (re->n_states > 0 && re->accepting[re->start_state]) { *s = start; *e = start; return 1; }
    return 0;
  }
  int32_t cur = re->start_state;
  int64_t pos = start, last = -1;
  if (cur < re->n_states && re->accepting[cur]) last = pos;
  while (pos < n) {
    int32_t nx = re->trans[(int64_t)cur * 256 + t[pos]];
    if (nx == -1) break;
    cur = nx;
    ++pos;
    if (re->accepting[cur]) last = pos;
  }
  if (pos == n && cur < re->n_states && re->accepting[cur]) last = pos;
  if (last != -1) {
    if (re->has_end_anchor && last != n) return 0;
    *s = start; *e = last; return 1;
  }
  return 0;
}

/* DFAEngine._optimized_simd_search, dfa.mojo:2199-2253 */
static int optimized_simd_search(const orx_regex *re, const uint8_t *t, int64_t n, int64_t start,
                                 int64_t *s, int64_t *e) {
  if (!re->has_matcher) return 0;
  int64_t pos = start;
  if (re->scan_eligible) {
    while (pos < n) {
      int64_t mp = find_first_nibble_match(re, t, pos, n);
      if (mp == -1) return 0;
      int64_t ml = count_consecutive_matches(re, t, mp, n);
      if (ml > 0) {
        int64_t me = mp + ml;
        if (re->has_end_anchor && me != n) { pos = me; continue; }
        *s = mp; *e = me; return 1;
      }
      pos = mp + 1;
    }
    return 0;
  }
  while (pos < n) {
    int64_t fp = find_first_nibble_match(re, t, pos, n);
    if (fp == -1) return 0;
    if (try_match_at_position(re, t, n, fp, 0, s, e)) return 1;
    pos = fp + 1;
  }
  return 0;
}

/* DFAEngine.match_next, dfa.mojo:1874-1903 */
static int dfa_match_next(const orx_regex *re, const uint8_t *t, int64_t n, int64_t start, int64_t *s, int64_t *e) {
  if (re->has_start_anchor) {
    if (start == 0) return try_match_at_position(re, t, n, 0, 0, s, e);
    return 0;
  }
  if (re->has_matcher && !re->has_end_anchor) return optimized_simd_search(re, t, n, start, s, e);
  for (int64_t p = start; p <= n; ++p)
    if (try_match_at_position(re, t, n, p, 0, s, e)) return 1;
  return 0;
}

/* DFAEngine.match_first, dfa.mojo:1851-1872 */
static int dfa_match_first(const orx_regex *re, const uint8_t *t, int64_t n, int64_t start, int64_t *s, int64_t *e) {
  if (re->has_start_anchor && start > 0) return 0;
  return try_match_at_position(re, t, n, start, 1, s, e);
}

/* DFAEngine.match_all, dfa.mojo:2028-2130 */
static void dfa_match_all(const orx_regex *re, const uint8_t *t, int64_t n, orx_spans *out) {
  int64_t s, e;
  if (re->has_start_anchor || re->has_end_anchor) { /* :2046-2050 */
    if (dfa_match_next(re, t, n, 0, &s, &e)) spans_push(out, s, e);
    return;
  }
  int64_t pos = 0;
  if (re->is_pure_literal) { /* :2062-2071 */
    int64_t m = re->literal_len;
    while (pos <= n - m) {
      int64_t hit = simd_search(re->literal, m, t, n, pos);
      if (hit == -1) break;
      spans_push(out, hit, hit + m);
      pos = hit + m;
    }
    return;
  }
  if (re->has_matcher && re->n_states > 0) {
    if (re->scan_eligible) { /* :2075-2094 */
      while (pos < n) {
        int64_t mp = find_first_nibble_match(re, t, pos, n);
        if (mp == -1) break;
        int64_t ml = count_consecutive_matches(re, t, mp, n);
        if (ml > 0) { spans_push(out, mp, mp + ml); pos = mp + ml; }
        else pos = mp + 1;
      }
      return;
    }
    while (pos < n) { /* :2097-2115 */
      int64_t np = find_first_nibble_match(re, t, pos, n);
      if (np == -1) break;
      pos = np;
      if (try_match_at_position(re, t, n, pos, 0, &s, &e)) {
        spans_push(out, s, e);
        if (e == s) pos += 1; else pos = e;
      } else {
        pos += 1;
      }
    }
    return;
  }
  while (pos <= n) { /* :2118-2128 */
    if (try_match_at_position(re, t, n, pos, 0, &s, &e)) {
      spans_push(out, s, e);
      if (e == s) pos += 1; else pos = e;
    } else {
      pos += 1;
    }
  }
}

/* HybridMatcher.match_first, matcher.mojo:733-753 */
int orx_match_first(const orx_regex *re, const uint8_t *t, int64_t n, int64_t start, int64_t *out) {
  if (re->is_wildcard_any) {
    if (start <= n) { out[0] = start; out[1] = n; return 1; }
    return 0;
  }
  return dfa_match_first(re, t, n, start, &out[0], &out[1]);
}

/* HybridMatcher.match_next, matcher.mojo:755-802 */
int orx_match_next(const orx_regex *re, const uint8_t *t, int64_t n, int64_t start, int64_t *out) {
  if (re->is_wildcard_any) {
    if (start <= n) { out[0] = start; out[1] = n; return 1; }
    return 0;
  }
  if (re->is_exact_literal && !re->has_anchors_info && re->exact_len > 0) { /* :768-781 */
    if (start >= n) return 0;
    int64_t pos = str_find(t, n, re->exact_literal, re->exact_len, start);
    if (pos != -1 && pos + re->exact_len <= n) { out[0] = pos; out[1] = pos + re->exact_len; return 1; }
    return 0;
  }
  if (re->prefilter && !re->has_anchors_info) { /* :784-796, prefilter.mojo:418-428 */
    if (start >= n) return 0;
    int64_t c = str_find(t, n, re->prefilter, re->prefilter_len, start);
    if (c == -1) return 0;
    return dfa_match_next(re, t, n, c, &out[0], &out[1]);
  }
  return dfa_match_next(re, t, n, start, &out[0], &out[1]);
}

/* The reference has no batch API: its callers loop `search(pattern, line)` over the documents
 * (matcher.mojo:1325-1338).  This is that loop, compiled, so that the CPU baseline is not a
 * Python-overhead measurement.  Document i = t[i*stride, min((i+1)*stride, n)); out[2i..2i+1] relative to
 * the document, (-1,-1) = None. */
void orx_search_batch(const orx_regex *re, const uint8_t *t, int64_t n, int64_t stride, int64_t n_docs, int64_t *out) {
  for (int64_t i = 0; i < n_docs; ++i) {
    const int64_t lo = i * stride, hi = lo + stride < n ? lo + stride : n;
    int64_t se[2];
    if (hi >= lo && orx_match_next(re, t + lo, hi - lo, 0, se)) { out[2 * i] = se[0]; out[2 * i + 1] = se[1]; }
    else { out[2 * i] = -1; out[2 * i + 1] = -1; }
  }
}

/* HybridMatcher._match_all_required_byte, matcher.mojo:864-898 */
static void match_all_required_byte(const orx_regex *re, const uint8_t *t, int64_t n, orx_spans *out) {
  uint8_t needle = (uint8_t)re->required_byte;
  int64_t pos = 0;
  while (pos < n) {
    int64_t hit = simd_find_byte(t, pos, n, needle);
    if (hit == -1) break;
    int64_t start = hit;
    while (start > 0 && re->lookup[t[start - 1]]) --start;
    int64_t s, e;
    if (dfa_match_first(re, t, n, start, &s, &e) && e > hit) {
      spans_push(out, s, e);
      pos = e;
      if (pos <= hit) pos = hit + 1;
    } else {
      pos = hit + 1;
    }
  }
}

/* HybridMatcher.match_all, matcher.mojo:804-862.  Returns malloc'ed pairs via *out_v. */
int64_t orx_match_all(const orx_regex *re, const uint8_t *t, int64_t n, int64_t **out_v) {
  orx_spans sp = {0, 0, 0};
  if (n >= 1024) { /* capacity hint text_len >> 7 (dfa.mojo:2041-2043); not observable */
    sp.cap = (n >> 7) > 8 ? (n >> 7) : 8;
    sp.v = (int64_t *)malloc((size_t)sp.cap * 2 * sizeof(int64_t));
  }
  if (re->is_wildcard_any) {
    spans_push(&sp, 0, n);
  } else if (re->is_exact_literal && !re->has_anchors_info) { /* :815-847 (overlapping, start=pos+1) */
    int64_t m = re->exact_len;
    if (m > 0 && m <= n) {
      int64_t start = 0, max_start = n - m;
      while (start <= max_start) {
        int64_t pos = str_find(t, n, re->exact_literal, m, start);
        if (pos == -1) break;
        if (pos + m <= n) spans_push(&sp, pos, pos + m);
        start = pos + 1;
        if (start > max_start) break;
      }
    }
  } else if (re->required_byte >= 0) {
    match_all_required_byte(re, t, n, &sp);
  } else {
    dfa_match_all(re, t, n, &sp);
  }
  *out_v = sp.v;
  return sp.n;
}

/* ---- sub: matcher.mojo:1593-1621 (_apply_template_fixed), 1679-1854 ----------------- */
typedef struct {
  uint8_t *v;
  int64_t n, cap;
} orx_buf;

static void buf_append(orx_buf *b, const uint8_t *p, int64_t m) {
  if (m <= 0) return;
  if (b->n + m > b->cap) {
    while (b->n + m > b->cap) b->cap = b->cap ? b->cap * 2 : 64;
    b->v = (uint8_t *)realloc(b->v, (size_t)b->cap);
  }
  memcpy(b->v + b->n, p, (size_t)m);
  b->n += m;
}

typedef struct {
  int32_t group_ref, start, length;
} orx_seg;

static void apply_template_fixed(orx_buf *out, const orx_seg *tpl, int32_t ntpl, const uint8_t *repl,
                                 const uint8_t *t, int64_t ms, const orx_regex *re) {
  for (int32_t i = 0; i < ntpl; ++i) {
    if (tpl[i].group_ref > 0 && tpl[i].group_ref <= re->fixed_num_groups)
      buf_append(out, t + ms + re->fixed_offsets[tpl[i].group_ref], re->fixed_widths[tpl[i].group_ref]);
    else
      buf_append(out, repl + tpl[i].start, tpl[i].length);
  }
}

/* _sub_impl_with_repl.  Returns 0 ok; -2 = general capture-group path (NFA engine,
 * matcher.mojo:1783-1821) which is out of scope for the oracle. */
int orx_sub(const orx_regex *re, const uint8_t *repl, int64_t repl_len, int use_groups, const orx_seg *tpl,
            int32_t ntpl, const uint8_t *t, int64_t n, int64_t count, uint8_t **out_v, int64_t *out_n) {
  orx_buf out = {0, 0, 0};
  *out_v = 0;
  *out_n = 0;
  if (n == 0) return 0; /* :1697-1698 */
  out.cap = n + 64;
  out.v = (uint8_t *)malloc((size_t)out.cap);
  int64_t pos = 0, reps = 0, m[2];
  if (use_groups) {
    if (re->fixed_total_width < 0) { free(out.v); return -2; }
    int64_t tw = re->fixed_total_width;
    if (re->fixed_concat && n == tw) { /* :1726-1744 */
      int all_digits = 1;
      for (int64_t i = 0; i < tw; ++i)
        if (t[i] < '0' || t[i] > '9') { all_digits = 0; break; }
      if (all_digits) apply_template_fixed(&out, tpl, ntpl, repl, t, 0, re);
      else buf_append(&out, t, n);
      *out_v = out.v; *out_n = out.n;
      return 0;
    }
    while (pos <= n) { /* :1747-1781 */
      if (!orx_match_next(re, t, n, pos, m)) break;
      if (m[0] > pos) buf_append(&out, t + pos, m[0] - pos);
      apply_template_fixed(&out, tpl, ntpl, repl, t, m[0], re);
      ++reps;
      if (m[1] == m[0]) {
        if (pos < n) buf_append(&out, t + pos, 1);
        pos = m[1] + 1;
      } else {
        pos = m[1];
      }
      if (count > 0 && reps >= count) break;
    }
  } else {
    while (pos <= n) { /* :1824-1847 */
      if (!orx_match_next(re, t, n, pos, m)) break;
      if (m[0] > pos) buf_append(&out, t + pos, m[0] - pos);
      buf_append(&out, repl, repl_len);
      ++reps;
      if (m[1] == m[0]) {
        if (pos < n) buf_append(&out, t + pos, 1);
        pos = m[1] + 1;
      } else {
        pos = m[1];
      }
      if (count > 0 && reps >= count) break;
    }
  }
  if (pos < n) buf_append(&out, t + pos, n - pos); /* :1849-1852 */
  *out_v = out.v;
  *out_n = out.n;
  return 0;
}

/* ---- split: matcher.mojo:1357-1393.  Pieces are returned as (bytes, offsets[P+1]).
 * Returns the number of pieces, or -3 when a separator starts before the previous one
 * ended (text[prev_end:start] with prev_end > start; Mojo slice behaviour unknown,
 * SURVEY appendix B-3) -- the oracle refuses instead of guessing. */
int64_t orx_split(const orx_regex *re, const uint8_t *t, int64_t n, int64_t maxsplit, uint8_t **out_bytes,
                  int64_t **out_offsets) {
  int64_t *mv = 0;
  int64_t nm = orx_match_all(re, t, n, &mv);
  orx_buf bytes = {0, 0, 0};
  int64_t *offs = (int64_t *)malloc((size_t)(nm + 2) * sizeof(int64_t));
  int64_t np = 0, prev_end = 0, splits = 0;
  offs[0] = 0;
  for (int64_t i = 0; i < nm; ++i) {
    if (maxsplit != 0 && splits >= maxsplit) break;
    int64_t s = mv[2 * i], e = mv[2 * i + 1];
    if (s < prev_end) { free(mv); free(offs); free(bytes.v); return -3; }
    buf_append(&bytes, t + prev_end, s - prev_end);
    offs[++np] = bytes.n;
    prev_end = e;
    ++splits;
  }
  buf_append(&bytes, t + prev_end, n - prev_end);
  offs[++np] = bytes.n;
  free(mv);
  *out_bytes = bytes.v;
  *out_offsets = offs;
  return np;
}
