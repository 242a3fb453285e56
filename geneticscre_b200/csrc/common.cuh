// Shared device/host definitions of the join engine (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gcre {

// one top-K candidate appended by the join kernels; `key` is an order-preserving integer image of the f64 score
struct Cand {
  unsigned long long key;
  uint32_t idx;
  uint32_t loc;
  int32_t cases;
  int32_t ctrls;
};

// order-preserving map f64 -> u64 (larger double <=> larger key); NaN must be filtered by the caller
__host__ __device__ __forceinline__ unsigned long long score_key(double x) {
#ifdef __CUDA_ARCH__
  unsigned long long b = (unsigned long long)__double_as_longlong(x);
#else
  unsigned long long b;
  memcpy(&b, &x, 8);
#endif
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__host__ __device__ __forceinline__ double key_score(unsigned long long k) {
  unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)b);
#else
  double x;
  memcpy(&x, &b, 8);
  return x;
#endif
}

// Everything a join kernel launch needs (passed by value).
struct JoinParams {
  // operands: row-major bitset matrices, row stride = Wp * M words, method-2 rows are [pos Wp | neg Wp]
  const uint64_t* p0;
  const uint64_t* p1;
  uint64_t* pres;          // nullptr unless the joined rows are kept
  int Wp;                  // words per half-row (even)
  int n_cases;
  // join index (UidRelSet, src/gcre.h:49-90)
  const int32_t* count;    // [U]
  const uint32_t* location;  // [U]
  const unsigned long long* prefix;   // [U+1] running sum of counts == flattened pair index of each uid's first pair
  const unsigned long long* res_idx;  // [U]   uid.path_idx: first result row of each uid
  uint32_t n_uids;
  const int32_t* signs;
  int path_length;
  unsigned long long pair_begin, pair_end;  // flattened pair range of this launch
  // permutation masks
  const uint64_t* pm;      // dense kernels: word-major [Wp][Ip]
  const uint32_t* pt;      // sparse kernels: patient-major [n][Iw] (bit r of word r/32)
  int Ip;                  // padded permutation count (multiple of the perm tile)
  int Iw;                  // 32-bit words per patient row of pt
  int n_perm_tiles;
  // value tables in anti-diagonal-major layout: entry (total, c) at total*(total+1)/2 + c
  const double* diagD;     // vt[c][total-c]                       (true scores)
  const float* diagF;      // (float) vt[c][total-c]               (method-1 permutation look-ups)
  const double* diagDM;    // max(vt[c][total-c], vt[total-c][c])  (method-2 permutation look-ups)
  const float* diagFM;     // diagDM rounded UP to f32: upper bounds for the "can any permutation beat the maxima" test
  // outputs
  int* perm_max;           // float bit patterns, >= +0.0f, merged with integer atomicMax
  Cand* cand;
  unsigned* cand_count;
  unsigned cand_cap;
  unsigned long long thr_key;  // append candidates whose key is strictly greater
  // Self-tightening threshold (sparse kernels, top_k <= 64).  slots[n_slots = top_k] hold, per hash bucket of (idx, loc),
  // the largest key appended so far; once every bucket is filled there are top_k distinct pairs with key >= min(slots),
  // so a pair strictly below that minimum cannot be in the top K.  dyn_thr caches the minimum (0 = not yet known).
  unsigned long long* dyn_thr;
  unsigned long long* slots;
  int n_slots;
  unsigned* max_total;     // atomicMax of per-half carrier totals of kept rows
};

// bits [0, n_cases) of the patient vector, word k  (src/join_base.cpp:50-54)
__device__ __forceinline__ uint64_t case_mask_word(int k, int n_cases) {
  long long lo = (long long)k * 64;
  if (lo + 64 <= n_cases) return ~0ull;
  if (lo >= n_cases) return 0ull;
  return (1ull << (n_cases - lo)) - 1ull;
}

// UidRelSet::need_flip (src/gcre.h:71-81)
__device__ __forceinline__ bool need_flip(int path_length, const int32_t* __restrict__ signs, unsigned idx, unsigned loc) {
  int sign;
  if (path_length > 3) sign = signs[idx];
  else if (path_length < 3) sign = signs[loc];
  else sign = (signs[idx] + signs[loc] == 0) ? -1 : 1;
  return sign == 1;
}

// flattened pair index -> upstream row: largest u with prefix[u] <= p  (prefix has n_uids + 1 entries)
__device__ __forceinline__ uint32_t find_uid(const unsigned long long* __restrict__ prefix, uint32_t n_uids, unsigned long long p) {
  uint32_t lo = 0, hi = n_uids;  // invariant: prefix[lo] <= p < prefix[hi]
  while (hi - lo > 1) {
    uint32_t mid = lo + ((hi - lo) >> 1);
    if (prefix[mid] <= p) lo = mid; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ size_t diag_base(unsigned total) { return (size_t)total * ((size_t)total + 1) / 2; }

}  // namespace gcre
