// Drop-in replacement for the reference's src/gcre_types.h: the same type names, fields and check helpers, so that
// src/wrapper.cpp, src/RcppExports.cpp and test/harness.cpp of geneticsCRE compile unchanged against this directory.
// (Reference: src/gcre_types.h:11-76.)  Authored for the B200 engine; nothing here computes on the CPU.
#ifndef GCRE_TYPES_H
#define GCRE_TYPES_H

#include <cstddef>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <vector>

// widest entity counts (src/gcre_types.h:11-13)
using st_path_count = uint64_t;
using st_uids_size = uint32_t;
using st_pathset_size = uint32_t;

using vec_i = std::vector<int>;
using vec_d = std::vector<double>;
using vec_u64 = std::vector<uint64_t>;
using vec2d_d = std::vector<std::vector<double>>;
using vec2d_f = std::vector<std::vector<float>>;
using vec2d_i = std::vector<std::vector<int>>;
using vec2d_u64 = std::vector<std::vector<uint64_t>>;
using vec2d_u16 = std::vector<std::vector<uint16_t>>;
using vec2d_i8 = std::vector<std::vector<int8_t>>;

const uint64_t bit_one_ul = 1;
const uint64_t bit_zero_ul = 0;

enum class Method { method1 = 1, method2 = 2 };

// One scored (upstream row, partner row) pair (src/gcre_types.h:32-43).  src/trg are ROW indices into the join's
// operands, not gene uids.  operator< is reversed so a std::priority_queue<Score> is a min-heap on score.
class Score {
 public:
  double score = -std::numeric_limits<double>::infinity();
  int src = -1;
  int trg = -1;
  int cases = 0;
  int ctrls = 0;
  Score() {}
  Score(double score_, int src_, int trg_, int cases_, int ctrls_) : score(score_), src(src_), trg(trg_), cases(cases_), ctrls(ctrls_) {}
  friend bool operator<(Score a, Score b) { return a.score > b.score; }
};

// Result of one join (src/gcre_types.h:45-48): top-K scores ascending, per-permutation maxima (float-rounded).
struct joined_res {
  std::vector<Score> scores;
  vec_d permuted_scores;
};

// One upstream row of a join index (src/gcre_types.h:50-56).
struct uid_ref {
  int src;
  int trg;
  int count;
  st_pathset_size location;
  st_path_count path_idx;
};

// Assertions of the reference throw these exact types/messages (src/gcre_types.h:58-76).
inline void check_true(bool condition) {
  if (!condition) throw std::logic_error("assertion");
}
inline void check_equal(size_t one, size_t two) {
  if (one != two) throw std::logic_error("assertion");
}
inline void check_index(long value, size_t size) {
  if (value < 0 || (size_t)value >= size) throw std::out_of_range("assertion");
}
inline void check_range(long value, long min, long max) {
  if (value < min || value > max) throw std::out_of_range("assertion");
}

#endif
