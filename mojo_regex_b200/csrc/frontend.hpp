// Host-side front end of the product: pattern -> the reference's hand-built automaton.
//
// In the real integration the Mojo front end (lexer/parser/optimizer/dfa.mojo builders)
// stays in place and hands its DFAEngine to mrx_dfa_pack().  This C++ front end exists so
// that the library is usable (and testable, and benchable) without a Mojo toolchain; it
// reproduces the reference's routing and table construction for every pattern the
// reference serves from its DFAEngine.  Reference citations are file:line relative to
// /root/reference/src/regex/.
#pragma once
#include <array>
#include <cstdint>
#include <string>
#include <vector>

namespace mrx {

struct RawDfa {  // DFAEngine, dfa.mojo:257-294
  std::vector<std::array<int32_t, 256>> trans;  // -1 = no transition (dfa.mojo:215-254)
  std::vector<uint8_t> accepting;
  int start_state = 0;
  bool has_start_anchor = false, has_end_anchor = false;
  bool is_pure_literal = false;
  std::string literal;  // literal_pattern
  bool has_matcher = false;
  std::array<uint8_t, 256> lookup{};  // _simd_char_matcher.lookup_table
  bool scan_eligible = false;
};

struct RegexMeta {  // HybridMatcher + CompiledRegex fields, matcher.mojo:535-565, 940-962
  bool is_wildcard_any = false;
  bool is_exact_literal = false;
  bool has_anchors_info = false;  // literal_info.has_anchors (false when analysis was skipped)
  std::string exact_literal;      // literal_info.best_literal
  std::string prefilter;          // MemchrPrefilter literal, empty = none
  int required_byte = -1;
  int fixed_total_width = -1, fixed_num_groups = 0;
  bool fixed_concat = false;
  int fixed_offsets[10] = {0}, fixed_widths[10] = {0};
};

// Returns MRX_OK / MRX_E_PARSE / MRX_E_UNSUPPORTED and fills `err` on failure.
int compile_frontend(const std::string &pattern, RawDfa &dfa, RegexMeta &meta, std::string &err);

}  // namespace mrx
