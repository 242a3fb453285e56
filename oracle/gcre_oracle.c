/*
 * TEST INFRASTRUCTURE ONLY.  CPU restatement (plain scalar C, 64-bit indexing, one shared value table) of the
 * geneticsCRE path-join + permutation-scoring path.  It exists to check the CUDA path; nothing in the product
 * (geneticscre_b200/, include/) may include, link or call it.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline leg of bench.py use it.
 *
 * Parity status: PINNED.  The reference ships no golden vectors, so this restatement is
 * pinned against the reference itself: tests/test_oracle.py runs it side by side with the unmodified reference
 * classes compiled from /root/reference (oracle/build_ref.sh -> oracle/_ref/libgcre_ref_*.so) on seeded inputs,
 * against the hand-derived known-answer case of SURVEY.md App. C, and against the tests/golden fixtures which were
 * produced by the reference build (tests/golden/make_golden.py).
 *
 * Each function cites the reference lines it follows (paths relative to /root/reference).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef struct {
  double score;
  int32_t src, trg, cases, ctrls;
} oracle_score;

static inline int popc64(uint64_t x) { return __builtin_popcountll(x); }

/* src/gcre_paths.h:56-78 -- PathSet::load: any non-zero entry sets bit c%64 of word c/64 in the FIRST W words
 * (the "pos" half for method 2); rows are vlen = W*method words. */
void oracle_pack_rows_i32(const int32_t* data, int64_t rows, int64_t cols, int W, int vlen, uint64_t* out) {
  memset(out, 0, (size_t)rows * vlen * sizeof(uint64_t));
  for (int64_t r = 0; r < rows; r++)
    for (int64_t c = 0; c < cols; c++)
      if (data[r * cols + c] != 0) out[r * vlen + c / 64] |= 1ull << (c % 64);
}

/* src/gcre_paths.h:82-92 -- PathSet::select: row gather. */
void oracle_select_rows(const uint64_t* src, int vlen, const int32_t* idx, int64_t n, uint64_t* out) {
  for (int64_t k = 0; k < n; k++) memcpy(out + k * vlen, src + (int64_t)idx[k] * vlen, vlen * sizeof(uint64_t));
}

/* src/join_base.cpp:50-54 (case mask = bits [0,num_cases)) and :85-125 (setPermutedCases):
 * mask_r = case_mask XOR flip_r, flip_r bit c set iff perm[r][c] != 1; rows reused cyclically when fewer rows
 * than iterations are supplied (:116-123), surplus rows ignored (:89-90).
 * Output layout here is perm-major [iters][W] (the reference stores word-major [W][iters_padded]). */
int oracle_perm_masks(const int32_t* perm, int64_t rows, int64_t cols, int num_cases, int W, int iters, uint64_t* out) {
  memset(out, 0, (size_t)iters * W * sizeof(uint64_t));
  if (iters > 0 && rows == 0) return -1; /* reference divides by zero here */
  int64_t have = rows < iters ? rows : iters;
  for (int64_t r = 0; r < have; r++) {
    uint64_t* m = out + r * W;
    for (int k = 0; k < num_cases; k++) m[k / 64] |= 1ull << (k % 64);
    for (int64_t c = 0; c < cols; c++)
      if (perm[r * cols + c] != 1) m[c / 64] ^= 1ull << (c % 64);
  }
  for (int64_t r = have; r < iters; r++) memcpy(out + r * W, out + (r % rows) * W, W * sizeof(uint64_t));
  return 0;
}

/* src/join_base.cpp:62-80 -- setValueTable pads the R table to (n+1)x(n+1) with -1.0. */
static inline double vt_at(const double* vt, int rows, int cols, int64_t r, int64_t c) {
  if (r < rows && c < cols) return vt[r * cols + c];
  return -1.0;
}

/* src/methods.h:110-118 -- value_table_max[r][c] = std::max(vt[r][c], vt[c][r]) over the padded square table.
 * std::max(a,b) returns (a < b) ? b : a. */
static inline double vtmax_at(const double* vt, int rows, int cols, int64_t r, int64_t c) {
  double a = vt_at(vt, rows, cols, r, c), b = vt_at(vt, rows, cols, c, r);
  return (a < b) ? b : a;
}

/* src/gcre.h:71-81 -- UidRelSet::need_flip */
static inline int need_flip(int path_length, const int32_t* signs, int64_t idx, int64_t loc) {
  int sign = 0;
  if (path_length > 3) sign = signs[idx];
  else if (path_length < 3) sign = signs[loc];
  else sign = (signs[idx] + signs[loc] == 0) ? -1 : 1;
  return sign == 1;
}

/* Deterministic top-K: K largest scores, ties broken by smaller (idx, loc) (the reference's
 * own membership among ties is heap- and thread-dependent).  Pairs are visited in ascending (idx, loc) order, so
 * "insert iff not full or score strictly greater than the current minimum" realises exactly that rule.
 * src/methods.h:90-94 (push iff score > heap min, pop while size > top_k). */
typedef struct {
  oracle_score* e; /* sorted descending by (score, then insertion order) */
  int n, k;
} topk_t;

static void topk_push(topk_t* t, double score, int idx, int loc, int cases, int ctrls) {
  if (!(score > -INFINITY)) return; /* never beats the -inf sentinel (strict >), NaN never pushed */
  if (t->n == t->k && !(score > t->e[t->n - 1].score)) return;
  int pos = t->n < t->k ? t->n : t->k - 1;
  while (pos > 0 && t->e[pos - 1].score < score) {
    t->e[pos] = t->e[pos - 1];
    pos--;
  }
  t->e[pos].score = score;
  t->e[pos].src = idx;
  t->e[pos].trg = loc;
  t->e[pos].cases = cases;
  t->e[pos].ctrls = ctrls;
  if (t->n < t->k) t->n++;
}

/*
 * The join: src/join_base.cpp:189-264 (loops over uids / partner ranges, optional result rows at path_idx),
 * src/methods.h:58-105 (method 1) and :130-232, :253-264 (method 2), result formatting src/join_base.cpp:138-154.
 *
 *  method        1 or 2;  W = words per half-row; rows of paths0/paths1/paths_res are vlen = W*method words
 *  masks         perm-major [iters][W]
 *  vt            row-major rows x cols (R layout [cases][ctrls]); padded with -1.0 outside
 *  paths_res     NULL (score only) or count_total x vlen, written at the running sum of counts
 *  out_scores    capacity top_k; ascending score order; a leading -inf sentinel {src=trg=-1} when fewer than
 *                top_k real entries exist (src/join_base.cpp:192-194)
 *  out_perm      iters floats: (float) max(+0, max over pairs p_score)
 * returns number of entries in out_scores, or -1 on a failed pre-check (src/join_base.cpp:196-200).
 */
int oracle_join(int method, int num_cases, int num_ctrls, int W, int iters, const uint64_t* masks, const double* vt,
                int vt_rows, int vt_cols, int path_length, int64_t n_uids, const int32_t* count,
                const uint32_t* location, const int32_t* signs, const uint64_t* paths0, int64_t n0,
                const uint64_t* paths1, int64_t n1, uint64_t* paths_res, int top_k, oracle_score* out_scores,
                float* out_perm) {
  const int vlen = W * method;
  if (n_uids != n0) return -1;
  for (int64_t u = 0; u < n_uids; u++)
    if (count[u] > 0 && (int64_t)location[u] + count[u] - 1 >= n1) return -1;

  uint64_t* case_mask = (uint64_t*)calloc(W > 0 ? W : 1, sizeof(uint64_t));
  for (int k = 0; k < num_cases; k++) case_mask[k / 64] |= 1ull << (k % 64);

  uint32_t* pc = (uint32_t*)malloc(sizeof(uint32_t) * 2 * (iters > 0 ? iters : 1));
  uint64_t* joined = (uint64_t*)malloc(sizeof(uint64_t) * (vlen > 0 ? vlen : 1));
  for (int r = 0; r < iters; r++) out_perm[r] = 0.0f;

  topk_t tk;
  tk.k = top_k > 0 ? top_k : 1;
  tk.n = 0;
  tk.e = (oracle_score*)malloc(sizeof(oracle_score) * tk.k);

  int64_t path_idx = 0;
  for (int64_t idx = 0; idx < n_uids; idx++) {
    const uint64_t* p0 = paths0 + idx * vlen;
    for (int64_t j = 0; j < count[idx]; j++) {
      const int64_t loc = (int64_t)location[idx] + j;
      const uint64_t* p1 = paths1 + loc * vlen;
      if (method == 1) {
        /* src/methods.h:66-88 */
        int cases = 0, ctrls = 0;
        memset(pc, 0, sizeof(uint32_t) * iters);
        for (int k = 0; k < W; k++) {
          uint64_t jw = p0[k] | p1[k];
          joined[k] = jw;
          if (jw == 0) continue;
          cases += popc64(jw & case_mask[k]);
          ctrls += popc64(jw & ~case_mask[k]);
          for (int r = 0; r < iters; r++) pc[r] += popc64(jw & masks[(int64_t)r * W + k]);
        }
        /* :90-94 */
        topk_push(&tk, vt_at(vt, vt_rows, vt_cols, cases, ctrls), (int)idx, (int)loc, cases, ctrls);
        /* :96-103 */
        int total = cases + ctrls;
        for (int r = 0; r < iters; r++) {
          double p = vt_at(vt, vt_rows, vt_cols, pc[r], total - (int64_t)pc[r]);
          if (p > out_perm[r]) out_perm[r] = (float)p;
        }
      } else {
        /* src/methods.h:137-145: halves of the downstream operand are routed by need_flip */
        const int flip = need_flip(path_length, signs, idx, loc);
        const uint64_t* pos0 = p0;
        const uint64_t* neg0 = p0 + W;
        const uint64_t* pos1 = flip ? p1 : p1 + W;
        const uint64_t* neg1 = flip ? p1 + W : p1;
        uint32_t* pcp = pc;         /* perm_case_pos */
        uint32_t* pnp = pc + iters; /* perm_ctrl_pos: |N & mask_r| */
        memset(pc, 0, sizeof(uint32_t) * 2 * iters);
        uint32_t case_pos = 0, case_neg = 0, ctrl_pos = 0, ctrl_neg = 0, total_pos = 0, total_neg = 0;
        /* :162-212 */
        for (int k = 0; k < W; k++) {
          uint64_t bp = pos0[k] | pos1[k], bn = neg0[k] | neg1[k];
          joined[k] = bp;
          joined[W + k] = bn;
          if (bp == 0 && bn == 0) continue;
          uint64_t m = case_mask[k];
          total_pos += popc64(bp);
          total_neg += popc64(bn);
          case_pos += popc64(bp & m);
          case_neg += popc64(bn & ~m);
          ctrl_pos += popc64(bn & m);
          ctrl_neg += popc64(bp & ~m);
          for (int r = 0; r < iters; r++) {
            uint64_t pm = masks[(int64_t)r * W + k];
            pcp[r] += popc64(bp & pm);
            pnp[r] += popc64(bn & pm);
          }
        }
        /* :253-264 keep_score */
        double score = vt_at(vt, vt_rows, vt_cols, case_pos, ctrl_neg) + vt_at(vt, vt_rows, vt_cols, case_neg, ctrl_pos);
        topk_push(&tk, score, (int)idx, (int)loc, (int)(case_pos + case_neg), (int)(ctrl_pos + ctrl_neg));
        /* :220-230 */
        for (int r = 0; r < iters; r++) {
          int64_t perm_case_neg = (int64_t)total_neg - pnp[r];
          int64_t perm_ctrl_neg = (int64_t)total_pos - pcp[r];
          double p = vtmax_at(vt, vt_rows, vt_cols, pcp[r], perm_ctrl_neg) + vtmax_at(vt, vt_rows, vt_cols, perm_case_neg, pnp[r]);
          if (p > out_perm[r]) out_perm[r] = (float)p;
        }
      }
      /* src/join_base.cpp:246-249 */
      if (paths_res) memcpy(paths_res + path_idx * vlen, joined, vlen * sizeof(uint64_t));
      path_idx++;
    }
  }

  /* src/join_base.cpp:138-154: ascending order; sentinel survives when fewer than top_k real entries */
  int n_out = 0;
  if (tk.n < tk.k) {
    out_scores[n_out].score = -INFINITY;
    out_scores[n_out].src = -1;
    out_scores[n_out].trg = -1;
    out_scores[n_out].cases = 0;
    out_scores[n_out].ctrls = 0;
    n_out++;
  }
  for (int k = tk.n - 1; k >= 0; k--) out_scores[n_out++] = tk.e[k];
  free(tk.e);
  free(pc);
  free(joined);
  free(case_mask);
  return n_out;
}

/* Score of one joined row recomputed from scratch -- independent restatement used to validate tied top-K entries
 * (R/CheckResults.R:50-73: method 1 VT[cases+1,controls+1]; method 2 VT[cp+1,cn'+1] + VT[cn+1,cp'+1]). */
double oracle_row_score(int method, int num_cases, int W, const uint64_t* row, const double* vt, int vt_rows, int vt_cols,
                        int* cases_out, int* ctrls_out) {
  int a = 0, b = 0, c = 0, d = 0;
  for (int k = 0; k < W; k++) {
    uint64_t cm = 0;
    for (int bit = 0; bit < 64; bit++)
      if (k * 64 + bit < num_cases) cm |= 1ull << bit;
    a += popc64(row[k] & cm);
    b += popc64(row[k] & ~cm);
    if (method == 2) {
      c += popc64(row[W + k] & ~cm); /* case_neg */
      d += popc64(row[W + k] & cm);  /* ctrl_pos */
    }
  }
  if (method == 1) {
    *cases_out = a;
    *ctrls_out = b;
    return vt_at(vt, vt_rows, vt_cols, a, b);
  }
  *cases_out = a + c;
  *ctrls_out = b + d;
  return vt_at(vt, vt_rows, vt_cols, a, b) + vt_at(vt, vt_rows, vt_cols, c, d);
}
