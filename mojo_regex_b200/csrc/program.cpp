// Blob pack / load and the scan planner.  Host only.
#include "program.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <queue>

#include "../../include/mojoregex_cuda.h"

namespace mrx {

bool pack_blob(const RawDfa &dfa, const RegexMeta &meta, std::string &blob, std::string &err) {
  const int ns = (int)dfa.trans.size();
  if (ns > kMaxStates) {
    err = "automaton has " + std::to_string(ns) + " states; the packed format holds " + std::to_string(kMaxStates);
    return false;
  }
  // byte classes: two bytes share a class iff every state sends them to the same target
  std::array<uint8_t, 256> cls{};
  int ncls = 0;
  {
    std::map<std::vector<int32_t>, int> seen;
    for (int b = 0; b < 256; ++b) {
      std::vector<int32_t> col(ns);
      for (int s = 0; s < ns; ++s) col[s] = dfa.trans[s][b];
      auto it = seen.find(col);
      if (it == seen.end()) it = seen.emplace(col, ncls++).first;
      cls[b] = (uint8_t)it->second;
    }
  }
  if (ns == 0) ncls = 0;
  BlobHeader h{};
  h.magic = kBlobMagic;
  h.version = kBlobVersion;
  h.flags = (dfa.has_start_anchor ? BF_START_ANCHOR : 0) | (dfa.has_end_anchor ? BF_END_ANCHOR : 0) |
            (dfa.is_pure_literal ? BF_PURE_LITERAL : 0) | (dfa.has_matcher ? BF_HAS_MATCHER : 0) |
            (dfa.scan_eligible ? BF_SCAN_ELIGIBLE : 0) | (meta.is_wildcard_any ? BF_WILDCARD_ANY : 0) |
            (meta.is_exact_literal ? BF_EXACT_LITERAL : 0) | (meta.has_anchors_info ? BF_HAS_ANCHORS_INFO : 0) |
            (!meta.prefilter.empty() ? BF_HAS_PREFILTER : 0);
  h.n_states = ns;
  h.n_classes = ncls;
  h.start_state = dfa.start_state;
  h.required_byte = meta.required_byte;
  h.fixed_total_width = meta.fixed_total_width;
  h.fixed_num_groups = meta.fixed_num_groups;
  h.fixed_concat = meta.fixed_concat ? 1 : 0;
  std::memcpy(h.fixed_offsets, meta.fixed_offsets, sizeof h.fixed_offsets);
  std::memcpy(h.fixed_widths, meta.fixed_widths, sizeof h.fixed_widths);
  h.literal_len = (uint32_t)dfa.literal.size();
  h.exact_len = (uint32_t)meta.exact_literal.size();
  h.prefilter_len = (uint32_t)meta.prefilter.size();
  blob.assign((const char *)&h, sizeof h);
  blob.append((const char *)cls.data(), 256);
  blob.append((const char *)dfa.lookup.data(), 256);
  blob.append((const char *)dfa.accepting.data(), ns);
  std::string tr((size_t)ns * ncls, (char)kDead);
  for (int s = 0; s < ns; ++s)
    for (int b = 0; b < 256; ++b) {
      int32_t t = dfa.trans[s][b];
      if (t < -1 || t >= ns) { err = "transition target out of range"; return false; }
      tr[(size_t)s * ncls + cls[b]] = (char)(t < 0 ? kDead : t);
    }
  blob += tr;
  blob += dfa.literal;
  blob += meta.exact_literal;
  blob += meta.prefilter;
  uint32_t total = (uint32_t)blob.size();
  std::memcpy(&blob[offsetof(BlobHeader, total_len)], &total, 4);
  return true;
}

namespace {

// ---- planner helpers ------------------------------------------------------------------------
void choose_filter(Plan &pl) {
  int cnt = 0, lo = -1, hi = -1, ranges = 0;
  bool in = false;
  for (int b = 0; b < 256; ++b) {
    if (pl.cand[b]) {
      ++cnt;
      if (!in) { ++ranges; in = true; if (lo < 0) lo = b; }
      hi = b;
    } else {
      in = false;
    }
  }
  if (cnt == 256) { pl.filt = FILT_ALL; return; }
  if (cnt >= 1 && cnt <= 4) {
    pl.filt = FILT_NEEDLES;
    pl.n_needles = 0;
    for (int b = 0; b < 256; ++b) if (pl.cand[b]) pl.needles[pl.n_needles++] = (uint8_t)b;
    for (int i = pl.n_needles; i < 4; ++i) pl.needles[i] = pl.needles[0];
    return;
  }
  if (ranges == 1) { pl.filt = FILT_RANGE; pl.range_lo = (uint8_t)lo; pl.range_hi = (uint8_t)hi; return; }
  pl.filt = FILT_LUT;
}

std::vector<uint8_t> can_reach_accept(const Program &p) {
  std::vector<uint8_t> ok(p.n_states, 0);
  for (int s = 0; s < p.n_states; ++s) ok[s] = p.accepting[s];
  bool changed = true;
  while (changed) {
    changed = false;
    for (int s = 0; s < p.n_states; ++s) {
      if (ok[s]) continue;
      for (int c = 0; c < p.n_classes; ++c) {
        int t = p.trans[(size_t)s * p.n_classes + c];
        if (t != kDead && ok[t]) { ok[s] = 1; changed = true; break; }
      }
    }
  }
  return ok;
}

// Can two successful attempts overlap (p < p' < A(p), both succeed)?  Conservative product search
// over (older thread state, younger thread state, older accepted after p', younger accepted).
bool attempts_can_overlap(const Program &p, const Plan &pl) {
  const int ns = p.n_states, nc = p.n_classes, D = ns;  // D = dead
  if (ns == 0) return false;
  std::vector<int> cls_rep(nc, -1);  // class -> is there a candidate byte in it
  std::vector<uint8_t> cls_cand(nc, 0);
  for (int b = 0; b < 256; ++b) if (pl.cand[b]) cls_cand[p.byte_class[b]] = 1;
  auto step = [&](int s, int c) { return s == D ? D : (p.trans[(size_t)s * nc + c] == kDead ? D : (int)p.trans[(size_t)s * nc + c]); };
  // R1: states reachable from start by >= 1 byte
  std::vector<uint8_t> r1(ns, 0);
  {
    std::queue<int> q;
    for (int c = 0; c < nc; ++c) { int t = step(p.start_state, c); if (t != D && !r1[t]) { r1[t] = 1; q.push(t); } }
    while (!q.empty()) {
      int s = q.front(); q.pop();
      for (int c = 0; c < nc; ++c) { int t = step(s, c); if (t != D && !r1[t]) { r1[t] = 1; q.push(t); } }
    }
  }
  auto idx = [&](int o, int y, int oa, int ya) { return ((o * (ns + 1) + y) * 2 + oa) * 2 + ya; };
  std::vector<uint8_t> seen((size_t)(ns + 1) * (ns + 1) * 4, 0);
  std::queue<std::array<int, 4>> q;
  auto push = [&](int o, int y, int oa, int ya) -> bool {
    if (o != D && p.accepting[o]) oa = 1;
    if (y != D && p.accepting[y]) ya = 1;
    if (oa && ya) return true;
    if (o == D && y == D) return false;
    if (o == D && !oa) return false;  // older died without accepting after p': cannot overlap
    if (y == D && !ya) return false;  // younger failed
    int k = idx(o, y, oa, ya);
    if (!seen[k]) { seen[k] = 1; q.push({o, y, oa, ya}); }
    return false;
  };
  for (int s = 0; s < ns; ++s) {
    if (!r1[s]) continue;
    for (int c = 0; c < nc; ++c) {
      if (!cls_cand[c]) continue;
      int y = step(p.start_state, c), o = step(s, c);
      if (y == D || o == D) continue;
      if (push(o, y, 0, 0)) return true;
    }
  }
  while (!q.empty()) {
    auto st = q.front(); q.pop();
    for (int c = 0; c < nc; ++c)
      if (push(step(st[0], c), step(st[1], c), st[2], st[3])) return true;
  }
  return false;
}

// REQUIRED_BYTE: can an accepted string contain a second needle after its first one?
bool second_needle_possible(const Program &p, int needle) {
  const int ns = p.n_states, nc = p.n_classes;
  const int ncl = p.byte_class[needle];
  // is the needle alone in its class?  if not, the class edge also covers other bytes; still
  // conservative to treat any edge on that class as a needle edge.
  std::vector<uint8_t> reach_ok = can_reach_accept(p);
  // phase 0: before first needle, phase 1: after first needle
  std::vector<uint8_t> seen(ns * 2, 0);
  std::queue<std::pair<int, int>> q;
  seen[p.start_state * 2] = 1;
  q.push({p.start_state, 0});
  while (!q.empty()) {
    auto [s, ph] = q.front(); q.pop();
    for (int c = 0; c < nc; ++c) {
      int t = p.trans[(size_t)s * nc + c];
      if (t == kDead) continue;
      int nph = ph;
      if (c == ncl) {
        if (ph == 1) { if (reach_ok[t]) return true; continue; }
        nph = 1;
      }
      if (!seen[t * 2 + nph]) { seen[t * 2 + nph] = 1; q.push({t, nph}); }
    }
  }
  return false;
}

bool literal_has_border(const std::string &s) {
  for (size_t k = 1; k < s.size(); ++k)
    if (s.compare(0, s.size() - k, s, k, s.size() - k) == 0) return true;
  return false;
}

// CharacterClassSIMD candidate set of find_first_nibble_match's vector body
// (simd_ops.mojo:62-86, 364-401, 565-648).
void simd_candidates(const std::array<uint8_t, 256> &lookup, std::array<uint8_t, 256> &cand) {
  int ranges = 0;
  bool in = false;
  for (int b = 0; b < 256; ++b) {
    if (lookup[b]) { if (!in) { ++ranges; in = true; } } else in = false;
  }
  if (ranges >= 1 && ranges <= 3) { cand = lookup; return; }
  uint8_t lo[16] = {0}, hi[16] = {0};
  int bucket = 0;
  for (int b = 0; b < 256; ++b)
    if (lookup[b]) {
      uint8_t bit = (uint8_t)(1u << (bucket & 7));
      lo[b & 15] |= bit;
      hi[b >> 4] |= bit;
      ++bucket;
    }
  for (int b = 0; b < 256; ++b) cand[b] = (lo[b & 15] & hi[b >> 4]) ? 1 : 0;
}

// Chain automaton over a single byte set with a contiguous accepting tail -> closed-form runs.
bool detect_uniform_run(const Program &p, const Plan &pl, int &mn, int &mx) {
  if (pl.shortcut || p.n_states < 2 || p.start_accepting()) return false;
  const int nc = p.n_classes;
  std::vector<uint8_t> xcls(nc, 0);
  for (int c = 0; c < nc; ++c) xcls[c] = p.trans[(size_t)p.start_state * nc + c] != kDead;
  for (int b = 0; b < 256; ++b)
    if ((pl.cand[b] != 0) != (xcls[p.byte_class[b]] != 0)) return false;  // candidates == live set
  int s = p.start_state, depth = 0, first_acc = -1;
  std::vector<uint8_t> visited(p.n_states, 0);
  while (true) {
    visited[s] = 1;
    int target = -1, live = 0, nx = 0;
    for (int c = 0; c < nc; ++c) {
      int t = p.trans[(size_t)s * nc + c];
      nx += xcls[c];
      if (t == kDead) continue;
      if (!xcls[c]) return false;  // edge on a byte outside the run class
      if (target != -1 && target != t) return false;
      target = t;
      ++live;
    }
    if (live == 0) break;          // terminal state: no outgoing edges at all
    if (live != nx) return false;  // only part of the class continues
    if (visited[target]) return false;  // loop
    s = target;
    ++depth;
    if (p.accepting[s]) { if (first_acc < 0) first_acc = depth; }
    else if (first_acc >= 0) return false;  // accepting set must be a contiguous tail
  }
  if (first_acc < 0 || depth < 1) return false;
  mn = first_acc;
  mx = depth;
  return true;
}

// Enumerate the (finite) language of the automaton from the start state.  Fails on cycles, on more than
// `max_strings` strings, on strings longer than 255 bytes, and when an accepting state can still reach
// another accept (then "longest accept" matters and the language is not prefix-free).
bool enumerate_language(const Program &p, std::vector<std::string> &out, size_t max_strings) {
  const int nc = p.n_classes;
  std::vector<std::vector<int>> class_bytes(nc);
  for (int b = 0; b < 256; ++b) class_bytes[p.byte_class[b]].push_back(b);
  std::vector<uint8_t> on_path(p.n_states, 0);
  std::string cur;
  bool ok = true;
  std::vector<uint8_t> reach = can_reach_accept(p);
  // iterative DFS with explicit recursion via lambda
  struct Rec {
    const Program &p; std::vector<std::vector<int>> &cb; std::vector<uint8_t> &on; std::string &cur;
    std::vector<std::string> &out; size_t maxs; bool &ok; std::vector<uint8_t> &reach;
    void go(int s) {
      if (!ok) return;
      if (on[s]) { ok = false; return; }
      if (p.accepting[s]) {
        for (int c = 0; c < p.n_classes; ++c) {
          int t = p.trans[(size_t)s * p.n_classes + c];
          if (t != kDead && reach[t]) { ok = false; return; }
        }
        if (cur.empty() || cur.size() > 255 || out.size() >= maxs) { ok = false; return; }
        out.push_back(cur);
        return;
      }
      on[s] = 1;
      for (int c = 0; c < p.n_classes && ok; ++c) {
        int t = p.trans[(size_t)s * p.n_classes + c];
        if (t == kDead || !reach[t]) continue;
        for (int b : cb[c]) {
          if (cur.size() >= 255) { ok = false; break; }
          cur.push_back((char)b);
          go(t);
          cur.pop_back();
          if (!ok) break;
        }
      }
      on[s] = 0;
    }
  } rec{p, class_bytes, on_path, cur, out, max_strings, ok, reach};
  rec.go(p.start_state);
  return ok && !out.empty();
}

// Aho-Corasick completion.  Fails (returns false) when a literal occurs inside / as a suffix of a
// proper prefix of another one (two occurrences could then overlap or nest) unless `allow_suffix` --
// for a single literal nesting is impossible.
bool build_dense(const std::vector<std::string> &lits, DenseTable &dt) {
  struct Node { int next[256]; int fail = 0; int out = 0; };
  std::vector<Node> nodes(1);
  std::fill(std::begin(nodes[0].next), std::end(nodes[0].next), -1);
  size_t maxlen = 0;
  for (const std::string &w : lits) {
    int cur = 0;
    for (unsigned char ch : w) {
      if (nodes[cur].next[ch] < 0) {
        nodes[cur].next[ch] = (int)nodes.size();
        nodes.emplace_back();
        std::fill(std::begin(nodes.back().next), std::end(nodes.back().next), -1);
      }
      cur = nodes[cur].next[ch];
      if (nodes[cur].out) return false;  // another literal is a proper prefix of this one
    }
    if (nodes[cur].out) continue;  // duplicate
    for (int b = 0; b < 256; ++b) if (nodes[cur].next[b] >= 0) return false;  // this literal is a prefix of another
    nodes[cur].out = (int)w.size();
    maxlen = std::max(maxlen, w.size());
  }
  if (nodes.size() > 120) return false;
  // BFS: failure links + complete transitions
  std::queue<int> q;
  for (int b = 0; b < 256; ++b) {
    int t = nodes[0].next[b];
    if (t < 0) nodes[0].next[b] = 0; else { nodes[t].fail = 0; q.push(t); }
  }
  while (!q.empty()) {
    int u = q.front(); q.pop();
    if (nodes[nodes[u].fail].out) return false;  // a literal ends inside another: occurrences nest
    for (int b = 0; b < 256; ++b) {
      int t = nodes[u].next[b];
      if (t < 0) nodes[u].next[b] = nodes[nodes[u].fail].next[b];
      else { nodes[t].fail = nodes[nodes[u].fail].next[b]; q.push(t); }
    }
  }
  // Row placement.  Rows are 512 bytes of uint16 entries; the entry for (row, byte) sits at byte offset
  // row + 2 * byte and holds the byte offset of the next row.  Rows are spaced 520 bytes apart so that
  // different rows map the same text byte to different shared-memory banks (a 512-byte spacing makes
  // every lane that is not in the root state collide with the root-state lanes).  Plain rows start at
  // multiples of 8; "a literal ends here" rows start at 8k + 4, so bit 2 of a row offset is the accept flag.
  // (Moving every non-root row another 16 banks away from the hot root row was measured: no gain.)
  std::vector<int> off(nodes.size(), -1);
  int n_plain = 0, n_out = 0;
  auto row_at = [](int k) { return 520 * k; };
  for (size_t i = 0; i < nodes.size(); ++i) if (!nodes[i].out) off[i] = row_at(n_plain++);
  for (size_t i = 0; i < nodes.size(); ++i) if (nodes[i].out) off[i] = row_at(n_plain + n_out++) + 4;
  const int table_bytes = ((row_at(n_plain + n_out) + 8) + 15) & ~15;
  if (table_bytes > 65536 || n_out > 32) return false;  // 32: len_lut is indexed by (offset >> 3) & 31, accept rows are consecutive
  dt.acc_base = n_plain;
  dt.n_states = (int)nodes.size();
  dt.max_len = (int)maxlen;
  dt.next.assign((size_t)table_bytes / 2, 0);
  dt.out_len.assign(n_out, 0);
  dt.len_lut.assign(32, 0);
  for (size_t i = 0; i < nodes.size(); ++i) {
    for (int b = 0; b < 256; ++b) dt.next[(size_t)off[i] / 2 + b] = (uint16_t)off[nodes[i].next[b]];
    if (nodes[i].out) {
      dt.out_len[(off[i] - 4) / 520 - n_plain] = (uint8_t)nodes[i].out;
      dt.len_lut[(off[i] >> 3) & 31] = (uint8_t)nodes[i].out;
    }
  }
  dt.valid = true;
  return true;
}

void try_dense(const Program &p, Plan &pl) {
  pl.dense = DenseTable();
  pl.sand_valid = false;
  if (p.flag(BF_START_ANCHOR) || p.flag(BF_END_ANCHOR)) return;
  std::vector<std::string> lits;
  if (pl.mode == MRX_MODE_LITERAL || pl.mode == MRX_MODE_EXACT_LITERAL) {
    if (pl.literal.empty() || pl.literal.size() > 255) return;
    lits.push_back(pl.literal);
  } else if (pl.mode == MRX_MODE_EVERY_POS && pl.direct && !pl.shortcut && !p.start_accepting()) {
    if (!enumerate_language(p, lits, 64)) return;
  } else {
    return;
  }
  DenseTable dt;
  if (build_dense(lits, dt) && dt.next.size() * 2 <= 24576) {
    pl.dense = dt;
    // the same set as a bit-parallel Shift-And automaton when it fits 32 state bits (distinct literals only:
    // build_dense has already refused prefixes and nesting)
    std::vector<std::string> uniq;
    for (const std::string &w : lits)
      if (std::find(uniq.begin(), uniq.end(), w) == uniq.end()) uniq.push_back(w);
    pl.sand_valid = sand::build(uniq, pl.sand, getenv("MRX_SAND_NO_TAIL") == nullptr);  // (the tests compile both forms)
  }
}

// Recognise, on the automaton the reference built, the shape
//   S0 --X--> S1,  S0 --c--> S2;   S1 --X--> S1,  S1 --c--> S2;   S2(acc) --Y--> S3;   S3(acc) --Y--> S3
// (dfa.mojo:699-807 for `X* c Y*`; S1 may coincide with S0, S3 with S2, X / Y may be empty), with
// c the required byte of the required-byte findall (matcher.mojo:864-898), the back-scan set L (first-class lookup)
// containing X and not c, and Y == X -- the preconditions of sepbits.hpp -- and fill the class table of k_split.
bool detect_separator_shape(const Program &p, sepb::Table &out) {
  const int c = p.required_byte, n = p.n_states, start = p.start_state;
  if (n <= 0 || c < 0 || c > 255 || p.accepting[start] || !p.flag(BF_HAS_MATCHER)) return false;
  if (p.flag(BF_START_ANCHOR) || p.flag(BF_END_ANCHOR)) return false;
  auto T = [&](int q, int b) { const int t = p.next(q, b); return t == kDead ? -1 : t; };
  const int s2 = T(start, c);
  if (s2 < 0 || !p.accepting[s2] || p.lookup[c]) return false;
  bool X[256] = {false}, Y[256] = {false};
  int s1 = -1;
  for (int b = 0; b < 256; ++b) {
    if (b == c) continue;
    const int t = T(start, b);
    if (t < 0) continue;
    if (s1 >= 0 && t != s1) return false;
    s1 = t;
    X[b] = true;
  }
  if (s1 >= 0 && s1 != start) {
    if (p.accepting[s1]) return false;
    for (int b = 0; b < 256; ++b) {
      const int t = T(s1, b);
      if (b == c) { if (t != s2) return false; }
      else if (X[b]) { if (t != s1) return false; }
      else if (t >= 0) return false;
    }
  }
  int s3 = -1;
  for (int b = 0; b < 256; ++b) {
    const int t = T(s2, b);
    if (t < 0) continue;
    if (b == c) return false;
    if (s3 >= 0 && t != s3) return false;
    s3 = t;
    Y[b] = true;
  }
  if (s3 >= 0 && s3 != s2) {
    if (!p.accepting[s3]) return false;
    for (int b = 0; b < 256; ++b) {
      const int t = T(s3, b);
      if (Y[b]) { if (t != s3) return false; }
      else if (t >= 0) return false;
    }
  }
  out = sepb::Table();
  for (int b = 0; b < 256; ++b) {
    if (X[b] != Y[b]) return false;
    if (X[b] && !p.lookup[b]) return false;
    out.T[b] = (b == c ? 1u : 0u) | (p.lookup[b] ? 1u << 8 : 0u) | ((p.lookup[b] && !X[b]) ? 1u << 16 : 0u);
  }
  return true;
}

// S0 --C--> S1(acc) --C--> S1 and nothing else (`\s+`, `(\d+)` as the LazyDFA builds them): an attempt at a class byte
// runs to the end of the class run, so the every-position loops (pikevm.mojo:757-803) return exactly the maximal runs
// -- the class-runs plan, without one attempt per byte of every run.
bool detect_class_plus(const Program &p, std::array<uint8_t, 256> &cls) {
  if (p.n_states != 2 || p.start_accepting() || p.flag(BF_START_ANCHOR) || p.flag(BF_END_ANCHOR) || p.flag(BF_HAS_MATCHER)) return false;
  const int s0 = p.start_state, s1 = 1 - s0;
  if (!p.accepting[s1]) return false;
  bool any = false;
  for (int b = 0; b < 256; ++b) {
    const int t0 = p.next(s0, b), t1 = p.next(s1, b);
    if (t0 == kDead && t1 == kDead) { cls[b] = 0; continue; }
    if (t0 != s1 || t1 != s1) return false;
    cls[b] = 1;
    any = true;
  }
  return any;
}

void plan_attempt_mode(const Program &p, Plan &pl, bool first_class) {
  pl.shortcut = p.flag(BF_HAS_MATCHER) && p.n_states > 0 && (p.flag(BF_SCAN_ELIGIBLE) || p.start_accepting());
  if (!first_class && detect_class_plus(p, pl.cand)) {
    pl.mode = MRX_MODE_CLASS_RUNS;
    pl.try_at_end = false;
    pl.direct = true;
    choose_filter(pl);
    return;
  }
  if (first_class) {
    pl.mode = MRX_MODE_FIRST_CLASS;
    pl.cand = p.lookup;
    pl.try_at_end = false;
    std::array<uint8_t, 256> simd{};
    simd_candidates(p.lookup, simd);
    for (int b = 0; b < 256; ++b) {
      if (simd[b] && !p.lookup[b]) {
        bool alive = p.n_states > 0 && p.next(p.start_state, b) != kDead;
        if (alive || p.start_accepting()) {
          pl.unsupported = true;
          pl.why = "first-class nibble filter has false positives that reach the automaton; the reference's "
                   "result then depends on the host SIMD width (simd_ops.mojo:117-133)";
          return;
        }
      }
    }
  } else {
    pl.mode = MRX_MODE_EVERY_POS;
    pl.try_at_end = true;
    for (int b = 0; b < 256; ++b)
      pl.cand[b] = (p.start_accepting() || p.flag(BF_END_ANCHOR) || pl.shortcut ||
                    (p.n_states > 0 && p.next(p.start_state, b) != kDead)) ? 1 : 0;
  }
  int mn, mx;
  if (!p.flag(BF_START_ANCHOR) && !p.flag(BF_END_ANCHOR) && detect_uniform_run(p, pl, mn, mx)) {
    pl.mode = MRX_MODE_UNIFORM_RUN;
    pl.run_min = mn;
    pl.run_max = mx;
    pl.direct = true;
    choose_filter(pl);
    return;
  }
  pl.direct = !pl.shortcut && !p.start_accepting() && !p.flag(BF_END_ANCHOR) && !attempts_can_overlap(p, pl);
  choose_filter(pl);
}

}  // namespace

// How common a byte is in ordinary text (higher = more common); only used to choose among required bytes.
static int byte_commonness(int b) {
  if (b >= 'a' && b <= 'z') return 6;
  if (b == ' ') return 7;
  if (b == '\n') return 5;
  if (b >= '0' && b <= '9') return 4;
  if (b >= 'A' && b <= 'Z') return 3;
  if (b == '.' || b == ',' || b == '-' || b == '_' || b == '/' || b == ':' || b == '\'' || b == '"' || b == '\t') return 2;
  if (b >= 0x21 && b < 0x7f) return 1;
  return 0;
}

// A byte c is required when no accepting state can be reached from the start state without consuming c.
// Decided on the automaton the kernels walk (which carries the reference's overwrite quirks), never on the
// pattern text.  Returns the rarest such byte, or -1 when there is none or only common ones (lower-case
// letters, blanks: nearly every document holds them, a filter would not pay).
static int required_byte(const Program &p) {
  const int n = p.n_states;
  if (n <= 0 || p.start_accepting()) return -1;
  int best = -1, best_rank = 1 << 30;
  std::vector<int> stack;
  std::vector<char> seen((size_t)n);
  for (int c = 0; c < 256; ++c) {
    std::fill(seen.begin(), seen.end(), 0);
    stack.assign(1, p.start_state);
    seen[(size_t)p.start_state] = 1;
    bool reach = false;
    while (!stack.empty() && !reach) {
      const int q = stack.back();
      stack.pop_back();
      for (int b = 0; b < 256 && !reach; ++b) {
        if (b == c) continue;
        const int nx = p.next(q, b);
        if (nx == kDead || seen[(size_t)nx]) continue;
        if (p.accepting[nx]) reach = true;
        seen[(size_t)nx] = 1;
        stack.push_back(nx);
      }
    }
    if (reach) continue;
    const int rank = byte_commonness(c) * 256 + c;
    if (rank < best_rank) { best_rank = rank; best = c; }
  }
  if (best >= 0 && byte_commonness(best) > 5) return -1;
  return best;
}

bool build_reverse(const Program &p, const Plan &pl, RevTable &rt) {
  rt = RevTable();
  if (pl.unsupported || pl.shortcut || p.n_states <= 0 || p.n_states > 32 || p.start_accepting()) return false;
  if (p.flag(BF_START_ANCHOR) || p.flag(BF_END_ANCHOR)) return false;
  if (pl.mode != MRX_MODE_FIRST_CLASS && pl.mode != MRX_MODE_EVERY_POS && pl.mode != MRX_MODE_UNIFORM_RUN) return false;
  const int n = p.n_states, s0 = p.start_state;
  constexpr int kMaxRev = 48;
  uint64_t acc = 0;
  for (int q = 0; q < n; ++q) if (p.accepting[q]) acc |= 1ull << q;
  std::vector<uint64_t> sets{acc};
  std::map<uint64_t, int> ids{{acc, 0}};
  std::vector<int> next;  // next[i * 256 + b] = set id
  for (size_t i = 0; i < sets.size(); ++i) {
    next.resize((i + 1) * 256);
    const uint64_t S = sets[i];
    for (int b = 0; b < 256; ++b) {
      uint64_t T = 0;
      for (int q = 0; q < n; ++q) {
        const int nx = p.next(q, b);
        if (nx != kDead && ((S >> nx) & 1)) T |= 1ull << q;
      }
      // "an attempt at this byte succeeds" must coincide with "the start state is in the new set": the
      // reference only attempts at candidate bytes, so a live first byte that is not a candidate cannot
      // be expressed as a row flag
      if (((T >> s0) & 1) && !pl.cand[b]) return false;
      T |= acc;
      auto it = ids.find(T);
      int id;
      if (it == ids.end()) {
        if ((int)sets.size() >= kMaxRev) return false;
        id = (int)sets.size();
        ids.emplace(T, id);
        sets.push_back(T);
      } else {
        id = it->second;
      }
      next[i * 256 + b] = id;
    }
  }
  auto rev_off = [&](int id) { return 520 * id + (((sets[id] >> s0) & 1) ? 4 : 0); };
  rt.n_rev = (int)sets.size();
  rt.rev.assign(((size_t)(520 * rt.n_rev + 8 + 15) & ~(size_t)15) / 2, 0);
  for (int i = 0; i < rt.n_rev; ++i)
    for (int b = 0; b < 256; ++b) rt.rev[(size_t)rev_off(i) / 2 + b] = (uint16_t)rev_off(next[(size_t)i * 256 + b]);
  build_forward_table(p, rt.fwd, rt.fwd_start);
  rt.req_byte = required_byte(p);
  rt.valid = true;
  return true;
}

void build_forward_table(const Program &p, std::vector<uint16_t> &fwd, uint16_t &start) {
  const int n = p.n_states;
  auto fwd_off = [&](int q) { return 520 * (q + 1) + (p.accepting[q] ? 4 : 0); };
  fwd.assign(((size_t)(520 * (n + 1) + 8 + 15) & ~(size_t)15) / 2, 0);  // row 0 = dead: every entry 0
  for (int q = 0; q < n; ++q)
    for (int b = 0; b < 256; ++b) {
      const int nx = p.next(q, b);
      fwd[(size_t)fwd_off(q) / 2 + b] = nx == kDead ? 0 : (uint16_t)fwd_off(nx);
    }
  start = (uint16_t)fwd_off(p.start_state);
}

namespace {

void make_plans(Program &p) {
  Plan &fa = p.findall, &se = p.search;
  // _try_match_at_position consults _try_match_simd first whenever the engine has a matcher and
  // (scan-eligible or accepting start state) -- an engine property, identical for every plan.
  fa.shortcut = se.shortcut =
      p.flag(BF_HAS_MATCHER) && p.n_states > 0 && (p.flag(BF_SCAN_ELIGIBLE) || p.start_accepting());
  // ---- match_all: matcher.mojo:804-862 then dfa.mojo:2028-2130
  if (p.flag(BF_WILDCARD_ANY)) {
    fa.mode = se.mode = MRX_MODE_WILDCARD_ANY;
    fa.direct = se.direct = true;
    return;
  }
  if (p.flag(BF_EXACT_LITERAL) && !p.flag(BF_HAS_ANCHORS_INFO)) {
    fa.mode = MRX_MODE_EXACT_LITERAL;
    fa.literal = p.exact;
    fa.direct = true;  // start = pos + 1: every occurrence is reported (matcher.mojo:843)
    if (!p.exact.empty()) fa.cand[(uint8_t)p.exact[0]] = 1;
    choose_filter(fa);
  } else if (p.required_byte >= 0) {
    fa.mode = MRX_MODE_REQUIRED_BYTE;
    fa.cand[p.required_byte] = 1;
    fa.direct = !fa.shortcut && !p.flag(BF_PURE_LITERAL) && !second_needle_possible(p, p.required_byte);
    choose_filter(fa);
    fa.sep_valid = fa.direct && detect_separator_shape(p, fa.sep);
  } else if (p.flag(BF_START_ANCHOR) || p.flag(BF_END_ANCHOR)) {
    fa.mode = MRX_MODE_ANCHORED;
    fa.direct = true;
  } else if (p.flag(BF_PURE_LITERAL)) {
    fa.mode = MRX_MODE_LITERAL;
    fa.literal = p.literal;
    fa.cand[(uint8_t)p.literal[0]] = 1;
    fa.direct = !literal_has_border(p.literal);
    choose_filter(fa);
  } else if (p.flag(BF_HAS_MATCHER) && p.n_states > 0 && p.flag(BF_SCAN_ELIGIBLE)) {
    fa.mode = MRX_MODE_CLASS_RUNS;
    fa.cand = p.lookup;
    fa.direct = true;
    choose_filter(fa);
  } else if (p.flag(BF_HAS_MATCHER) && p.n_states > 0) {
    plan_attempt_mode(p, fa, true);
  } else {
    plan_attempt_mode(p, fa, false);
  }
  // ---- match_next: matcher.mojo:755-802 then dfa.mojo:1874-1903
  if (p.flag(BF_EXACT_LITERAL) && !p.flag(BF_HAS_ANCHORS_INFO) && !p.exact.empty()) {
    se.mode = MRX_MODE_EXACT_LITERAL;
    se.literal = p.exact;
    se.cand[(uint8_t)p.exact[0]] = 1;
    se.direct = !literal_has_border(p.exact);
    choose_filter(se);
  } else if (p.flag(BF_START_ANCHOR)) {
    se.mode = MRX_MODE_ANCHORED;
    se.direct = true;
  } else if (p.flag(BF_HAS_MATCHER) && !p.flag(BF_END_ANCHOR)) {
    if (p.flag(BF_SCAN_ELIGIBLE)) {
      se.mode = MRX_MODE_CLASS_RUNS;
      se.cand = p.lookup;
      se.direct = true;
      choose_filter(se);
    } else {
      plan_attempt_mode(p, se, true);
    }
  } else if (p.flag(BF_PURE_LITERAL)) {
    se.mode = MRX_MODE_LITERAL;
    se.literal = p.literal;
    se.cand[(uint8_t)p.literal[0]] = 1;
    se.direct = !literal_has_border(p.literal);
    choose_filter(se);
  } else {
    plan_attempt_mode(p, se, false);
  }
  try_dense(p, fa);
  try_dense(p, se);
  build_reverse(p, se, se.rev);
  // ---- match_next through the MemchrPrefilter: matcher.mojo:784-796
  if (p.flag(BF_HAS_PREFILTER) && !p.flag(BF_HAS_ANCHORS_INFO) && !p.prefilter.empty() &&
      !(p.flag(BF_EXACT_LITERAL) && !p.exact.empty())) {
    // is the literal a prefix of every match?  (then "first match at or after the next literal" = "first match")
    bool prefix = p.n_states > 0 && !p.start_accepting();
    int st = p.start_state;
    for (size_t i = 0; prefix && i < p.prefilter.size(); ++i) {
      int live = 0, nxt = kDead;
      for (int b = 0; b < 256; ++b) {
        const int t = p.next(st, b);
        if (t != kDead) { ++live; if (b == (uint8_t)p.prefilter[i]) nxt = t; }
      }
      prefix = live == 1 && nxt != kDead && (i + 1 == p.prefilter.size() || !p.accepting[nxt]);
      st = nxt;
    }
    if (!prefix) {
      if (se.unsupported || !(se.mode == MRX_MODE_FIRST_CLASS || se.mode == MRX_MODE_EVERY_POS || se.mode == MRX_MODE_LITERAL)) {
        se.unsupported = true;
        se.why = "match_next through the MemchrPrefilter (matcher.mojo:784-796) on a plan that is not an attempt plan";
      } else {
        p.search_via_prefilter = true;
        Plan &pre = p.prefilter_plan;
        pre = Plan();
        pre.mode = MRX_MODE_EXACT_LITERAL;  // String.find semantics: every occurrence, overlapping ones included
        pre.literal = p.prefilter;
        pre.direct = true;
        pre.cand[(uint8_t)p.prefilter[0]] = 1;
        choose_filter(pre);
      }
    }
  }
}

}  // namespace

bool load_program(const uint8_t *blob, size_t n, Program &p, std::string &err) {
  if (!blob || n < sizeof(BlobHeader)) { err = "blob too short"; return false; }
  BlobHeader h;
  std::memcpy(&h, blob, sizeof h);
  if (h.magic != kBlobMagic || h.version != kBlobVersion) { err = "bad blob magic/version"; return false; }
  if (h.total_len != n) { err = "blob length mismatch"; return false; }
  if (h.n_states < 0 || h.n_states > kMaxStates || h.n_classes < 0 || h.n_classes > 256) { err = "bad state/class count"; return false; }
  size_t need = sizeof h + 512 + (size_t)h.n_states + (size_t)h.n_states * h.n_classes + h.literal_len + h.exact_len + h.prefilter_len;
  if (need != n) { err = "blob sections do not add up"; return false; }
  if (h.n_states > 0 && (h.start_state < 0 || h.start_state >= h.n_states)) { err = "bad start state"; return false; }
  if (h.required_byte < -1 || h.required_byte > 255) { err = "bad required byte"; return false; }
  p = Program();
  p.flags = h.flags;
  p.n_states = h.n_states;
  p.n_classes = h.n_classes;
  p.start_state = h.start_state;
  p.required_byte = h.required_byte;
  p.fixed_total_width = h.fixed_total_width;
  p.fixed_num_groups = h.fixed_num_groups;
  p.fixed_concat = h.fixed_concat;
  std::memcpy(p.fixed_offsets, h.fixed_offsets, sizeof p.fixed_offsets);
  std::memcpy(p.fixed_widths, h.fixed_widths, sizeof p.fixed_widths);
  if (p.fixed_num_groups < 0 || p.fixed_num_groups > 9) { err = "bad fixed-sub group count"; return false; }
  if (p.fixed_total_width >= 0) {  // group slices index the match: every one must lie inside it (k_assemble reads them)
    if (p.fixed_total_width > 255) { err = "fixed-sub match width above 255"; return false; }
    for (int gi = 0; gi <= p.fixed_num_groups; ++gi) {
      const int64_t o = p.fixed_offsets[gi], w = p.fixed_widths[gi];
      if (o < 0 || w < 0 || o + w > p.fixed_total_width) { err = "fixed-sub group slice outside the match"; return false; }
    }
  }
  const uint8_t *q = blob + sizeof h;
  std::memcpy(p.byte_class.data(), q, 256); q += 256;
  std::memcpy(p.lookup.data(), q, 256); q += 256;
  p.accepting.assign(q, q + h.n_states); q += h.n_states;
  p.trans.assign(q, q + (size_t)h.n_states * h.n_classes); q += (size_t)h.n_states * h.n_classes;
  p.literal.assign((const char *)q, h.literal_len); q += h.literal_len;
  p.exact.assign((const char *)q, h.exact_len); q += h.exact_len;
  p.prefilter.assign((const char *)q, h.prefilter_len);
  for (int b = 0; b < 256; ++b)
    if (h.n_states > 0 && p.byte_class[b] >= h.n_classes) { err = "byte class out of range"; return false; }
  for (uint8_t t : p.trans)
    if (t != kDead && t >= h.n_states) { err = "transition target out of range"; return false; }
  if (p.flag(BF_PURE_LITERAL) && p.literal.empty()) { err = "pure literal without literal bytes"; return false; }
  if (!p.flag(BF_WILDCARD_ANY) && h.n_states == 0) { err = "empty automaton"; return false; }
  p.blob.assign((const char *)blob, n);
  make_plans(p);
  return true;
}

}  // namespace mrx
