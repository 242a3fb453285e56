// Value-table generation on the device (reference: getValuesTable, R/Utils.R:137-159).
//
//   vt[x][i - x] = -log( sum of hypergeometric probabilities <= P(X = x) ),   X ~ Hypergeom(cases, ctrls, i draws)
// for every total i in [0, n] and every feasible x; infinities are replaced by (largest finite entry + 1).
//
// The R code is O(n * range^2) with sapply and is infeasible for n >= 50,000; the numpy restatement in synth.py takes
// minutes there.  Here one CTA owns a diagonal i:
//   1. log-probabilities from a log-factorial table (host lgamma, uploaded), in the SAME association order as
//      synth.make_value_table so that both select the same "probabilities <= own" sets: the comparison is done on the
//      log-probabilities, which are bit-identical on host and device; only exp() and the summation order differ (1e-15);
//   2. the pmf is unimodal: left of the mode it ascends, right of it it descends.  Prefix sums over the left part and
//      suffix sums over the right part (both add small terms first) + two binary searches give each two-sided sum in
//      O(log m) instead of O(m).
#pragma once
#include "common.cuh"

namespace gcre {

constexpr int VT_THREADS = 256;
// "probabilities <= own" is decided with this tolerance on the log-probabilities (= 1e-7 relative on the probabilities, the
// relErr of R's own fisher.test).  Exact ties between different outcomes are common (P(x+1)/P(x) = 1 has integer solutions,
// and every symmetric table has them); evaluated exactly, the reference's `prob_dist <= x` includes them, while its float
// comparison keeps or drops them by dhyper's last-bit rounding.  With the tolerance the table equals the exact-arithmetic
// evaluation of getValuesTable (tests/test_value_table.py checks it against big-integer combinatorics).
constexpr double VT_TIE_TOL = 1e-7;

__device__ __forceinline__ double vt_logp(const double* __restrict__ lf, int nc, int nt, int i, int x) {
  // (lchoose(nc, x) + lchoose(nt, i - x)) - lchoose(nc + nt, i), each lchoose(a, b) = (lf[a] - lf[b]) - lf[a - b]
  const double a = (lf[nc] - lf[x]) - lf[nc - x];
  const double b = (lf[nt] - lf[i - x]) - lf[nt - (i - x)];
  const double c = (lf[nc + nt] - lf[i]) - lf[nc + nt - i];
  return (a + b) - c;
}

// block-wide inclusive scan of one value per thread (result in every thread; `total` = sum over the block)
__device__ __forceinline__ double block_inclusive_scan(double v, double* s_warp, double& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += t;
  }
  if (lane == 31) s_warp[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double w = (lane < VT_THREADS / 32) ? s_warp[lane] : 0.0;
#pragma unroll
    for (int d = 1; d < VT_THREADS / 32; d <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w += t;
    }
    if (lane < VT_THREADS / 32) s_warp[lane] = w;
  }
  __syncthreads();
  const double before = warp ? s_warp[warp - 1] : 0.0;
  total = s_warp[VT_THREADS / 32 - 1];
  __syncthreads();
  return v + before;
}

// scratch: per CTA 3 * m_max doubles (logp, left prefix sums, right suffix sums)
__global__ void __launch_bounds__(VT_THREADS) value_table_kernel(const double* __restrict__ lf, int nc, int nt, double* __restrict__ vt,
                                                                 double* __restrict__ scratch, int m_max, unsigned long long* __restrict__ max_finite_key) {
  __shared__ double s_warp[VT_THREADS / 32];
  __shared__ int s_kmax;
  __shared__ unsigned long long s_best;
  const int n = nc + nt, cols = nt + 1;
  double* logp = scratch + (size_t)blockIdx.x * 3 * m_max;
  double* lsum = logp + m_max;   // lsum[k] = sum of p over left-part indices <= k
  double* rsum = lsum + m_max;   // rsum[k] = sum of p over right-part indices >= k
  unsigned long long best_key = 0;

  for (int i = blockIdx.x; i <= n; i += gridDim.x) {
    const int lo = max(0, i - nt), hi = min(i, nc), m = hi - lo + 1;
    // 1. log-probabilities and the position of the maximum (first maximal index)
    if (threadIdx.x == 0) s_best = 0;
    __syncthreads();
    for (int k = threadIdx.x; k < m; k += VT_THREADS) logp[k] = vt_logp(lf, nc, nt, i, lo + k);
    __syncthreads();
    unsigned long long mine = 0;
    for (int k = threadIdx.x; k < m; k += VT_THREADS) {
      // order by (logp, smaller index first): key = ordered(logp) in the high bits is not enough with 64-bit doubles,
      // so take the max of logp first and the smallest index attaining it second
      const unsigned long long key = score_key(logp[k]);
      mine = max(mine, key);
    }
    atomicMax(&s_best, mine);
    __syncthreads();
    if (threadIdx.x == 0) s_kmax = m;
    __syncthreads();
    for (int k = threadIdx.x; k < m; k += VT_THREADS)
      if (score_key(logp[k]) == s_best) atomicMin(&s_kmax, k);
    __syncthreads();
    const int kmax = s_kmax;  // left part = [0, kmax] (ascending), right part = (kmax, m) (descending)

    // 2. prefix sums over the left part, suffix sums over the right part (chunk per thread, then a block scan)
    {
      const int nl = kmax + 1, chunk = (nl + VT_THREADS - 1) / VT_THREADS;
      const int b = min(nl, (int)threadIdx.x * chunk), e = min(nl, b + chunk);
      double acc = 0.0;
      for (int k = b; k < e; k++) acc += exp(logp[k]);
      double total;
      const double incl = block_inclusive_scan(acc, s_warp, total);
      double run = incl - acc;
      for (int k = b; k < e; k++) {
        run += exp(logp[k]);
        lsum[k] = run;
      }
    }
    {
      const int nr = m - (kmax + 1), chunk = (nr + VT_THREADS - 1) / VT_THREADS;
      // walk the right part from its end (small probabilities first): position t counts from the end
      const int b = min(nr, (int)threadIdx.x * chunk), e = min(nr, b + chunk);
      double acc = 0.0;
      for (int t = b; t < e; t++) acc += exp(logp[m - 1 - t]);
      double total;
      const double incl = block_inclusive_scan(acc, s_warp, total);
      double run = incl - acc;
      for (int t = b; t < e; t++) {
        run += exp(logp[m - 1 - t]);
        rsum[m - 1 - t] = run;
      }
    }
    __syncthreads();

    // 3. two-sided sums by binary search, -log, store on the anti-diagonal of the R layout
    for (int k = threadIdx.x; k < m; k += VT_THREADS) {
      const double v = logp[k] + VT_TIE_TOL;
      // left part ascending: number of entries <= v
      int a = 0, bnd = kmax + 1;
      while (a < bnd) {
        const int mid = (a + bnd) >> 1;
        if (logp[mid] <= v) a = mid + 1; else bnd = mid;
      }
      double two = a > 0 ? lsum[a - 1] : 0.0;
      // right part descending: first index whose entry <= v
      int c = kmax + 1, d = m;
      while (c < d) {
        const int mid = (c + d) >> 1;
        if (logp[mid] <= v) d = mid; else c = mid + 1;
      }
      if (c < m) two += rsum[c];
      const double val = -log(two);
      const int x = lo + k;
      vt[(size_t)x * cols + (i - x)] = val;
      if (val == val && val < INFINITY) best_key = max(best_key, score_key(val));
    }
    __syncthreads();
  }
  if (best_key) atomicMax(max_finite_key, best_key);
}

// infinities -> (largest finite + 1); -0.0 -> +0.0  (R/Utils.R:156; synth.make_value_table)
__global__ void value_table_fixup_kernel(double* __restrict__ vt, size_t entries, const unsigned long long* __restrict__ max_finite_key) {
  const double repl = key_score(*max_finite_key) + 1.0;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < entries; k += (size_t)gridDim.x * blockDim.x) {
    const double v = vt[k];
    if (!(v == v) || v == INFINITY || v == -INFINITY) vt[k] = repl;
    else if (v == 0.0) vt[k] = 0.0;
  }
}

}  // namespace gcre
