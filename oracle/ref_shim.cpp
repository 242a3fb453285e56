// TEST INFRASTRUCTURE ONLY -- never linked into or called by the product path.
//
// Flat C wrapper around the *unmodified* reference classes (JoinExec / PathSet / UidRelSet,
// /root/reference/src/gcre.h, gcre_paths.h, join_base.cpp) so that tests and the CPU-baseline leg of
// bench.py can drive the real reference from Python (ctypes).  This file is ours; it is compiled by
// oracle/build_ref.sh together with the reference sources *where they lie* under /root/reference --
// nothing from the reference is copied into this repository.  Output goes to oracle/_ref/ (git-ignored).
#include "gcre.h"   // resolved with -I/root/reference/src at build time
#include <cstring>
#include <string>
#include <stdexcept>

static thread_local std::string g_err;

#define REF_TRY try {
#define REF_CATCH(ret) } catch (const std::exception& e) { g_err = e.what(); return ret; } catch (...) { g_err = "unknown"; return ret; }

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

const char* ref_simd_label() { return gs_instr_label.c_str(); }
int ref_vec_width() { return gs_vec_width; }

void* ref_exec_create(const char* method, int num_cases, int num_ctrls, int iters, int top_k, int nthreads) {
  REF_TRY
  JoinExec* exec = new JoinExec(std::string(method), num_cases, num_ctrls, iters);
  exec->top_k = top_k;
  exec->nthreads = nthreads;
  return exec;
  REF_CATCH(nullptr)
}

void ref_exec_destroy(void* h) { delete static_cast<JoinExec*>(h); }

int ref_exec_width_ul(void* h) { return static_cast<JoinExec*>(h)->width_ul; }
int ref_exec_iterations(void* h) { return static_cast<JoinExec*>(h)->iterations; }
void ref_exec_set_threads(void* h, int nthreads) { static_cast<JoinExec*>(h)->nthreads = nthreads; }
void ref_exec_set_top_k(void* h, int top_k) { static_cast<JoinExec*>(h)->top_k = top_k; }

// row-major rows x cols doubles
int ref_exec_set_value_table(void* h, const double* tbl, int rows, int cols) {
  REF_TRY
  vec2d_d table(rows, vec_d(cols));
  for (int r = 0; r < rows; r++)
    for (int c = 0; c < cols; c++) table[r][c] = tbl[(size_t)r * cols + c];
  static_cast<JoinExec*>(h)->setValueTable(table);
  return 0;
  REF_CATCH(-1)
}

// row-major rows x cols ints (1 = label kept)
int ref_exec_set_perms(void* h, const int* perms, int rows, int cols) {
  REF_TRY
  vec2d_i data(rows, vec_i(cols));
  for (int r = 0; r < rows; r++)
    for (int c = 0; c < cols; c++) data[r][c] = perms[(size_t)r * cols + c];
  static_cast<JoinExec*>(h)->setPermutedCases(data);
  return 0;
  REF_CATCH(-1)
}

void* ref_pathset_create(void* h, unsigned size) {
  REF_TRY
  return static_cast<JoinExec*>(h)->createPathSet(size).release();
  REF_CATCH(nullptr)
}

void ref_pathset_destroy(void* p) { delete static_cast<PathSet*>(p); }
unsigned ref_pathset_size(void* p) { return static_cast<PathSet*>(p)->size; }
int ref_pathset_vlen(void* p) { return static_cast<PathSet*>(p)->vlen; }
int ref_pathset_width_ul(void* p) { return static_cast<PathSet*>(p)->width_ul; }

int ref_pathset_load(void* p, const int* data, int rows, int cols) {
  REF_TRY
  vec2d_i d(rows, vec_i(cols));
  for (int r = 0; r < rows; r++)
    for (int c = 0; c < cols; c++) d[r][c] = data[(size_t)r * cols + c];
  static_cast<PathSet*>(p)->load(d);
  return 0;
  REF_CATCH(-1)
}

void* ref_pathset_select(void* p, const int* idx, int n) {
  REF_TRY
  std::vector<int> v(idx, idx + n);
  return static_cast<PathSet*>(p)->select(v).release();
  REF_CATCH(nullptr)
}

// copy all rows out: size x vlen words (reference padding included)
int ref_pathset_copy_out(void* p, uint64_t* out) {
  REF_TRY
  PathSet* ps = static_cast<PathSet*>(p);
  for (unsigned r = 0; r < ps->size; r++)
    std::memcpy(out + (size_t)r * ps->vlen, (*ps)[r], ps->vlen * sizeof(uint64_t));
  return 0;
  REF_CATCH(-1)
}

// raw row write (PathSet::set), used to feed packed rows without going through int matrices
int ref_pathset_set_row(void* p, unsigned row, const uint64_t* words) {
  REF_TRY
  static_cast<PathSet*>(p)->set(row, words);
  return 0;
  REF_CATCH(-1)
}

struct ref_score { double score; int src, trg, cases, ctrls; };

// uids given as parallel arrays; path_idx recomputed as the running sum of counts (test/harness.cpp:8-14)
// returns number of scores written (<= cap) or -1
int ref_join(void* h, int path_length, int n_uids, const int* src, const int* trg, const int* count,
             const unsigned* location, const int* signs, int n_signs, void* ps0, void* ps1, void* ps_res,
             ref_score* out_scores, int cap, double* out_perm) {
  REF_TRY
  std::vector<uid_ref> uids(n_uids);
  st_path_count total = 0;
  for (int k = 0; k < n_uids; k++) {
    uids[k].src = src[k];
    uids[k].trg = trg[k];
    uids[k].count = count[k];
    uids[k].location = location[k];
    uids[k].path_idx = total;
    total += count[k];
  }
  UidRelSet set(path_length, uids, std::vector<int>(signs, signs + n_signs));
  JoinExec* exec = static_cast<JoinExec*>(h);
  joined_res res = exec->join(set, *static_cast<PathSet*>(ps0), *static_cast<PathSet*>(ps1), *static_cast<PathSet*>(ps_res));
  int n = (int)res.scores.size();
  if (n > cap) n = cap;
  for (int k = 0; k < n; k++) {
    out_scores[k].score = res.scores[k].score;
    out_scores[k].src = res.scores[k].src;
    out_scores[k].trg = res.scores[k].trg;
    out_scores[k].cases = res.scores[k].cases;
    out_scores[k].ctrls = res.scores[k].ctrls;
  }
  for (size_t k = 0; k < res.permuted_scores.size(); k++) out_perm[k] = res.permuted_scores[k];
  return n;
  REF_CATCH(-1)
}

}  // extern "C"
