// Decorated p-values on the device (reference: computeDecoratedPvalue, R/DecoratedPvalue.R:198-304),
// in the exact limit of the reference's Monte-Carlo estimate.
//
// For a split of a reported path into a sub-path (pos1 / neg1 carrier vectors) and the gene it is extended by (pos2 / neg2),
// the reference redraws the carriers the gene ADDS uniformly among the patients the sub-path does not cover and counts how
// often the re-scored path reaches the real score.  Only the number of redrawn carriers that land on cases matters, and that
// number is hypergeometric, so the estimate converges to
//     p = sum over (cp, cn) of P_pos(cp) * P_neg(cn) * [score_of(cp, cn) >= score]
// which is what this kernel evaluates: the same AND / popcount / value-table look-up shape as the join, one CTA per split.
// (The stratified variant and the seeded Monte-Carlo mode stay host utilities in geneticscre_b200/decorated.py.)
#pragma once
#include "common.cuh"

namespace gcre {

constexpr int DEC_THREADS = 256;

struct DecoratedOut {
  double pvalue;
  double score;
  int32_t cases1, ctrls1, cases2, ctrls2;
};

__device__ __forceinline__ double dec_block_sum(double v, double* s_red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  __syncthreads();
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < DEC_THREADS / 32; w++) t += s_red[w];  // fixed order: every thread gets the same bits
  return t;
}

// value table read with the engine's convention for entries outside the supplied table (-1.0, src/join_base.cpp:72)
__device__ __forceinline__ double dec_vt(const double* __restrict__ vt, int rows, int cols, long long r, long long c) {
  return (r >= 0 && c >= 0 && r < rows && c < cols) ? vt[(size_t)r * cols + c] : -1.0;
}

// rows: [n_items][4][W64] packed pos1, neg1, pos2, neg2; scratch: [n_items][2][n + 1] doubles
template <int M>
__global__ void __launch_bounds__(DEC_THREADS) decorated_exact_kernel(const uint64_t* __restrict__ rows, uint32_t n_items, int W64, int n_cases, int n_ctrls,
                                                                     const double* __restrict__ vt, int vt_rows, int vt_cols,
                                                                     double* __restrict__ scratch, DecoratedOut* __restrict__ out) {
  __shared__ double s_red[DEC_THREADS / 32];
  __shared__ int s_cnt[8];
  const uint32_t item = blockIdx.x;
  if (item >= n_items) return;
  const int n = n_cases + n_ctrls;
  const uint64_t* pos1 = rows + (size_t)item * 4 * W64;
  const uint64_t* neg1 = pos1 + W64;
  const uint64_t* pos2 = neg1 + W64;
  const uint64_t* neg2 = pos2 + W64;
  if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  // R/DecoratedPvalue.R:206-225: the gene's carriers already in the sub-path do not count; eight AND + popcount sums
  int c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int k = threadIdx.x; k < W64; k += DEC_THREADS) {
    const uint64_t cm = case_mask_word(k, n_cases);
    const uint64_t valid = (k == W64 - 1 && (n & 63)) ? ((1ull << (n & 63)) - 1ull) : ~0ull;
    const uint64_t p1 = pos1[k] & valid, n1 = neg1[k] & valid;
    const uint64_t p2 = pos2[k] & valid & ~p1, n2 = neg2[k] & valid & ~n1;
    c[0] += __popcll(p1 & cm);   // case_pos1
    c[1] += __popcll(p1 & ~cm);  // control_pos1
    c[2] += __popcll(n1 & ~cm);  // case_neg1 (the negative part counts controls as its "cases")
    c[3] += __popcll(n1 & cm);   // control_neg1
    c[4] += __popcll(p2 & cm);   // case_pos2
    c[5] += __popcll(p2 & ~cm);  // control_pos2
    c[6] += __popcll(n2 & ~cm);  // case_neg2
    c[7] += __popcll(n2 & cm);   // control_neg2
  }
#pragma unroll
  for (int q = 0; q < 8; q++) {
    const int v = __reduce_add_sync(0xffffffffu, c[q]);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_cnt[q], v);
  }
  __syncthreads();
  const int case_pos1 = s_cnt[0], ctrl_pos1 = s_cnt[1], case_neg1 = s_cnt[2], ctrl_neg1 = s_cnt[3];
  const int case_pos2 = s_cnt[4], ctrl_pos2 = s_cnt[5], case_neg2 = s_cnt[6], ctrl_neg2 = s_cnt[7];
  const int k_pos = case_pos2 + ctrl_pos2, k_neg = case_neg2 + ctrl_neg2;
  // pools (R/DecoratedPvalue.R:232-233): everything outside the sub-path's part; "good" = a case (positive part) / a control (negative)
  const int good_pos = n_cases - case_pos1, bad_pos = n_ctrls - ctrl_pos1;
  const int good_neg = n_ctrls - case_neg1, bad_neg = n_cases - ctrl_neg1;

  auto score_of = [&](int cp, int cn) -> double {
    if (M == 1)  // R/DecoratedPvalue.R:226-227, 285-286
      return dec_vt(vt, vt_rows, vt_cols, (long long)case_pos1 + cp + case_neg1 + cn, (long long)ctrl_pos1 + (k_pos - cp) + ctrl_neg1 + (k_neg - cn));
    return dec_vt(vt, vt_rows, vt_cols, case_pos1 + cp, ctrl_pos1 + (k_pos - cp)) +
           dec_vt(vt, vt_rows, vt_cols, case_neg1 + cn, ctrl_neg1 + (k_neg - cn));  // :228-229, :287-288
  };
  const double score = score_of(case_pos2, case_neg2);

  // hypergeometric probabilities of the two independent counts over their supports
  const int lo_p = max(0, k_pos - bad_pos), hi_p = min(k_pos, good_pos), m_p = hi_p - lo_p + 1;
  const int lo_n = max(0, k_neg - bad_neg), hi_n = min(k_neg, good_neg), m_n = hi_n - lo_n + 1;
  double* pp = scratch + (size_t)item * 2 * (n + 1);
  double* pn = pp + (n + 1);
  // Unnormalised probabilities by the exact ratio recurrence P(x+1) / P(x) = (good - x)(k - x) / ((x + 1)(bad - k + x + 1)), started
  // at the mode and run outwards (one thread per distribution: <= n steps of two multiplications).  Relative error ~1e-16 per
  // step; log-gamma differences would lose ~n ln n * 1e-16 to cancellation.
  if (threadIdx.x == 0 || threadIdx.x == 32) {
    const bool neg = threadIdx.x == 32;
    const long long good = neg ? good_neg : good_pos, bad = neg ? bad_neg : bad_pos, k = neg ? k_neg : k_pos;
    const long long lo = neg ? lo_n : lo_p, hi = neg ? hi_n : hi_p;
    double* w = neg ? pn : pp;
    long long mode = (k + 1) * (good + 1) / (good + bad + 2);
    mode = mode < lo ? lo : (mode > hi ? hi : mode);
    w[mode - lo] = 1.0;
    for (long long x = mode; x < hi; x++) w[x + 1 - lo] = w[x - lo] * ((double)((good - x) * (k - x)) / (double)((x + 1) * (bad - k + x + 1)));
    for (long long x = mode; x > lo; x--) w[x - 1 - lo] = w[x - lo] * ((double)(x * (bad - k + x)) / (double)((good - x + 1) * (k - x + 1)));
  }
  __syncthreads();
  double sp = 0.0, sn = 0.0;
  for (int k = threadIdx.x; k < m_p; k += DEC_THREADS) sp += pp[k];
  for (int k = threadIdx.x; k < m_n; k += DEC_THREADS) sn += pn[k];
  sp = dec_block_sum(sp, s_red);
  sn = dec_block_sum(sn, s_red);
  __syncthreads();  // pp / pn are read by other threads below
  double acc = 0.0;
  const long long terms = (long long)m_p * m_n;
  for (long long q = threadIdx.x; q < terms; q += DEC_THREADS) {
    const int kp = (int)(q / m_n), kn = (int)(q % m_n);
    if (score_of(lo_p + kp, lo_n + kn) >= score) acc += pp[kp] * pn[kn];
  }
  acc = dec_block_sum(acc, s_red);
  if (threadIdx.x == 0) {
    DecoratedOut o;
    o.pvalue = acc / (sp * sn);
    o.score = score;
    o.cases1 = case_pos1 + case_neg1;
    o.ctrls1 = ctrl_pos1 + ctrl_neg1;
    o.cases2 = case_pos2 + case_neg2;
    o.ctrls2 = ctrl_pos2 + ctrl_neg2;
    out[item] = o;
  }
}

}  // namespace gcre
