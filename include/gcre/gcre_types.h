// Source-compatibility header of the B200 engine for code written against geneticsCRE's src/gcre_types.h
// (src/wrapper.cpp, src/RcppExports.cpp, test/harness.cpp, test/test.h): it provides the names those files use --
// Score, joined_res, uid_ref, Method, the st_* size types, the vec* aliases and the check_* helpers -- with the
// reference's observable behaviour (field names and order, exception types and messages).  Nothing here computes.
// Reference: src/gcre_types.h:11-76.
#ifndef GCRE_TYPES_H
#define GCRE_TYPES_H

#include <cstddef>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <vector>

// ---- scoring method selector (reference: enum class Method, values 1 and 2 are relied upon: vlen = width_ul * method)
enum class Method { method1 = 1, method2 = 2 };

// ---- integer widths of the entities the join handles
typedef uint64_t st_path_count;    // number of (upstream, partner) pairs / result rows of one join
typedef uint32_t st_uids_size;     // number of upstream rows
typedef uint32_t st_pathset_size;  // number of rows of a path set

// ---- container shorthands used by the callers
typedef std::vector<int> vec_i;
typedef std::vector<double> vec_d;
typedef std::vector<uint64_t> vec_u64;
typedef std::vector<vec_d> vec2d_d;
typedef std::vector<std::vector<float>> vec2d_f;
typedef std::vector<vec_i> vec2d_i;
typedef std::vector<vec_u64> vec2d_u64;
typedef std::vector<std::vector<uint16_t>> vec2d_u16;
typedef std::vector<std::vector<int8_t>> vec2d_i8;

static const uint64_t bit_zero_ul = 0, bit_one_ul = 1;

// ---- one scored pair.  `src` is the upstream ROW, `trg` the partner ROW of the join that produced it (not gene uids);
// a default-constructed Score is the "-infinity sentinel" that leads a result list with fewer than top_k real entries.
// The comparison is inverted on purpose: std::priority_queue<Score> then pops the LOWEST score first.
class Score {
 public:
  double score;
  int src, trg, cases, ctrls;

  Score() : score(-std::numeric_limits<double>::infinity()), src(-1), trg(-1), cases(0), ctrls(0) {}
  Score(double value, int upstream_row, int partner_row, int n_cases, int n_ctrls)
      : score(value), src(upstream_row), trg(partner_row), cases(n_cases), ctrls(n_ctrls) {}

  friend bool operator<(Score lhs, Score rhs) { return rhs.score < lhs.score; }
};

// ---- what JoinExec::join returns: the top-K scores in ascending order and, per permutation, the float-rounded maximum
// score over all pairs
struct joined_res {
  std::vector<Score> scores;
  vec_d permuted_scores;
};

// ---- one upstream row of a join index: its partners are rows [location, location + count) of the downstream set and its
// result rows start at path_idx.  Layout shared with gcre_uid_ref of the C ABI.
struct uid_ref {
  int src;
  int trg;
  int count;
  st_pathset_size location;
  st_path_count path_idx;
};

// ---- argument checks.  The reference reports every failed check as the string "assertion": logic_error for conditions
// and size mismatches, out_of_range for indices and ranges; callers (and Rcpp's BEGIN_RCPP) depend on those types.
namespace gcre_detail {
[[noreturn]] inline void fail_condition() { throw std::logic_error("assertion"); }
[[noreturn]] inline void fail_bounds() { throw std::out_of_range("assertion"); }
}  // namespace gcre_detail

inline void check_true(bool ok) {
  if (!ok) gcre_detail::fail_condition();
}
inline void check_equal(size_t lhs, size_t rhs) {
  if (lhs != rhs) gcre_detail::fail_condition();
}
inline void check_index(long index, size_t extent) {
  if (index < 0 || static_cast<size_t>(index) >= extent) gcre_detail::fail_bounds();
}
inline void check_range(long value, long lowest, long highest) {
  if (value < lowest || highest < value) gcre_detail::fail_bounds();
}

#endif  // GCRE_TYPES_H
