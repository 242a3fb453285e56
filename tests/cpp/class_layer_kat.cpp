// Known-answer test of the C++ drop-in class layer (include/gcre/*.h) itself, independent of the reference checkout:
// the hand-derived case of SURVEY.md App. C run through JoinExec / PathSet / UidRelSet exactly as src/wrapper.cpp uses
// them, plus the exception types the reference's checks throw (src/gcre_types.h:58-76).
// Built and run by tests/test_cpp_layer_gpu.py:  g++ -std=c++11 -Iinclude/gcre ... -lgcre_b200
#include <cmath>
#include <cstdio>
#include <stdexcept>

#include "gcre.h"
#include "util.h"

static int failures = 0;
#define EXPECT(cond)                                                    \
  do {                                                                  \
    if (!(cond)) {                                                      \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);     \
      failures++;                                                       \
    }                                                                   \
  } while (0)

static joined_res run(const char* method, int sign) {
  JoinExec exec(method, 4, 4, 1);
  exec.top_k = 2;
  vec2d_d table(9, vec_d(9));
  for (int i = 0; i < 9; i++)
    for (int j = 0; j < 9; j++) table[i][j] = 10.0 * i + j;
  exec.setValueTable(table);
  exec.setPermutedCases(vec2d_i(1, vec_i{1, 0, 1, 1, 0, 1, 1, 1}));
  TPathSet p0 = exec.createPathSet(1), p1 = exec.createPathSet(1), none = exec.createPathSet(0);
  p0->load(vec2d_i(1, vec_i{1, 1, 0, 0, 1, 0, 0, 0}));
  p1->load(vec2d_i(1, vec_i{0, 1, 1, 0, 0, 1, 0, 0}));
  vector<uid_ref> uids(1);
  uids[0].src = 7;
  uids[0].trg = 9;
  uids[0].count = 1;
  uids[0].location = 0;
  uids[0].path_idx = 0;
  UidRelSet set(2, uids, vector<int>{sign});
  Timer timer(exec, 2, set.count_total_paths());
  return exec.join(set, *p0, *p1, *none);
}

int main() {
  Timer::print_header();
  for (int m = 1; m <= 2; m++)
    for (int sign = -1; sign <= 1; sign += 2) {
      joined_res r = run(m == 1 ? "method1" : "method2", sign);
      EXPECT(r.scores.size() == 2);
      EXPECT(std::isinf(r.scores[0].score) && r.scores[0].score < 0 && r.scores[0].src == -1 && r.scores[0].trg == -1);
      const bool split = (m == 2 && sign == -1);  // the downstream gene goes to the negative half
      EXPECT(r.scores[1].score == (split ? 33.0 : 32.0));
      EXPECT(r.scores[1].src == 0 && r.scores[1].trg == 0);
      EXPECT(r.scores[1].cases == 3 && r.scores[1].ctrls == (split ? 3 : 2));
      EXPECT(r.permuted_scores.size() == 1 && r.permuted_scores[0] == (split ? 42.0 : 32.0));
    }
  {  // kept rows and row access
    JoinExec exec("method1", 4, 4, 1);
    exec.setValueTable(vec2d_d(9, vec_d(9, 1.0)));
    exec.setPermutedCases(vec2d_i(1, vec_i(8, 1)));
    TPathSet p0 = exec.createPathSet(1), p1 = exec.createPathSet(2), res = exec.createPathSet(2);
    p0->load(vec2d_i(1, vec_i{1, 0, 0, 0, 0, 0, 0, 0}));
    p1->load(vec2d_i{vec_i{0, 1, 0, 0, 0, 0, 0, 0}, vec_i{0, 0, 0, 0, 0, 0, 0, 1}});
    vector<uid_ref> uids(1);
    uids[0].src = uids[0].trg = 0;
    uids[0].count = 2;
    uids[0].location = 0;
    uids[0].path_idx = 0;
    exec.join(UidRelSet(2, uids, vector<int>{1, 1}), *p0, *p1, *res);
    EXPECT((*res)[0][0] == 0x03 && (*res)[1][0] == 0x81);
    TPathSet sel = res->select(vector<int>{1, 1, 0});
    EXPECT(sel->size == 3 && (*sel)[2][0] == 0x03);
    bool threw = false;
    try {
      (*res)[2];
    } catch (const std::out_of_range&) {
      threw = true;
    }
    EXPECT(threw);
    threw = false;
    try {  // uids.size() != paths0.size  -> std::logic_error("assertion") (src/join_base.cpp:196)
      exec.join(UidRelSet(2, uids, vector<int>{1, 1}), *p1, *p1, *res);
    } catch (const std::logic_error& e) {
      threw = std::string(e.what()) == "assertion";
    }
    EXPECT(threw);
  }
  std::printf(failures ? "class layer: %d FAILURES\n" : "class layer: OK\n", failures);
  return failures ? 1 : 0;
}
