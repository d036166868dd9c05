// Planner side of the next finder (see sep_core.hpp): recognise, on the automaton the reference built, the shape
//   S0 --X--> S1,  S0 --c--> S2;   S1 --X--> S1,  S1 --c--> S2;   S2(acc) --Y--> S3;   S3(acc) --Y--> S3
// (dfa.mojo:699-807 for `X* c Y*`; S1 may coincide with S0, S3 with S2, X / Y may be empty)
// and express the sets the way sep_core.hpp tests them.  NOT part of the library yet.  Host only.
//
// Conditions beyond the shape: c is the required byte of the required-byte findall (matcher.mojo:864-898); the
// back-scan set L (first-class lookup) contains X and not c; Y == X; L is one range of bytes below 128 plus at
// most two single bytes; at most two bytes of L are missing from X ("bad": the back-scan passes them, the
// automaton cannot consume them -- `\v` for `\s`).
#pragma once
#include <cstdint>
#include <vector>

#include "sep_core.hpp"

namespace mrx {
namespace sep {

// trans: n x 256, -1 = no transition
inline bool detect(const int32_t *trans, const uint8_t *accepting, int n, int start, const uint8_t *lookup, int c,
                   Classes &out) {
  if (n <= 0 || start < 0 || start >= n || c < 0 || c > 255 || accepting[start]) return false;
  auto T = [&](int q, int b) { return trans[(size_t)q * 256 + b]; };
  const int s2 = T(start, c);
  if (s2 < 0 || !accepting[s2] || lookup[c]) return false;
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
  if (s1 >= 0) {
    if (accepting[s1]) return false;
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
  if (s3 >= 0) {
    if (!accepting[s3]) return false;
    for (int b = 0; b < 256; ++b) {
      const int t = T(s3, b);
      if (Y[b]) { if (t != s3) return false; }
      else if (t >= 0) return false;
    }
  }
  std::vector<int> L, bad;
  for (int b = 0; b < 256; ++b) {
    if (X[b] != Y[b]) return false;
    if (X[b] && !lookup[b]) return false;
    if (lookup[b]) { L.push_back(b); if (!X[b]) bad.push_back(b); }
  }
  if (bad.size() > 2) return false;
  // L = longest run of consecutive bytes below 128 (if it has at least two members) + at most two others
  int best_lo = 1, best_hi = 0;
  for (size_t i = 0; i < L.size();) {
    size_t j = i;
    while (j + 1 < L.size() && L[j + 1] == L[j] + 1 && L[j + 1] < 128) ++j;
    if (L[i] < 128 && j > i && (int)(j - i) > best_hi - best_lo) { best_lo = L[i]; best_hi = L[j]; }
    i = j + 1;
  }
  out = Classes{};
  out.needle = (uint32_t)c * 0x01010101u;
  out.range_lo = (uint32_t)best_lo;
  out.range_hi = (uint32_t)best_hi;
  for (int b : L) {
    if (best_lo <= best_hi && b >= best_lo && b <= best_hi) continue;
    if (out.n_extra == 2) return false;
    out.extra[out.n_extra++] = (uint32_t)b * 0x01010101u;
  }
  for (int b : bad) out.bad[out.n_bad++] = (uint32_t)b * 0x01010101u;
  return true;
}

}  // namespace sep
}  // namespace mrx
