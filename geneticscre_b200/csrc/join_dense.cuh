// Dense join + permutation scoring kernel (sm_100a): AND + POPC of joined rows against word-major mask tiles.
//
// Replaces JoinMethod1::score_permute / JoinMethod2::score_permute (src/methods.h:58-105, 130-232) and the pair loop
// of JoinExec::join (src/join_base.cpp:230-250).  One CTA scores a tile of TP pairs x TI permutations:
//   * the joined rows (p0 | p1, method-2 halves routed by need_flip) are formed once per k-chunk into shared memory,
//     written to the result set when rows are kept, and their true case/control counts accumulated;
//   * the permutation mask chunk [KC words][TI perms] is staged with cp.async, double buffered;
//   * every thread owns an RP x RI register tile of (pair, perm) counters; mask words are read as 128-bit
//     conflict-free shared loads, joined words as 128-bit broadcasts;
//   * epilogue: anti-diagonal value-table look-ups, per-perm max in registers -> shared -> one atomicMax per
//     (CTA, perm) on the float bit image; per-pair true score -> thresholded top-K candidate append.
// Integer counts are exact under any summation order, so results are bit-identical to the reference.
#pragma once
#include "common.cuh"

namespace gcre {

namespace dense {
constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;
constexpr int RP = 4;            // pairs per thread
constexpr int RI = 4;            // perms per thread
constexpr int TP = WARPS * RP;   // 32 pairs per CTA
constexpr int TI = 32 * RI;      // 128 perms per CTA
constexpr int KC = 16;           // 64-bit words per k-chunk
}  // namespace dense

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int M, bool KEEP>
__global__ void __launch_bounds__(dense::THREADS, 2) join_dense_kernel(const JoinParams a) {
  using namespace dense;
  __shared__ __align__(16) uint64_t s_mask[2][KC][TI];     // 32 KB
  __shared__ __align__(16) uint64_t s_join[M][TP][KC];     // 4 / 8 KB
  __shared__ uint32_t s_idx[TP], s_loc[TP];
  __shared__ unsigned long long s_res_row[TP];
  __shared__ int s_flip[TP];
  __shared__ unsigned s_cnt[TP][M][2];                      // [pair][half][in case mask / outside]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int perm_tile = blockIdx.x % a.n_perm_tiles;
  const unsigned long long pair_tile = blockIdx.x / a.n_perm_tiles;
  const unsigned long long pair0 = a.pair_begin + pair_tile * TP;
  const int Wp = a.Wp;
  const int row_words = Wp * M;

  // ---- pair -> (idx, loc) for the tile ----
  if (tid < TP) {
    const unsigned long long p = pair0 + tid;
    uint32_t idx = 0xffffffffu, loc = 0;
    int flip = 1;
    unsigned long long rrow = 0;
    if (p < a.pair_end) {
      idx = find_uid(a.prefix, a.n_uids, p);
      const unsigned long long j = p - a.prefix[idx];
      loc = a.location[idx] + (uint32_t)j;
      if (M == 2) flip = need_flip(a.path_length, a.signs, idx, loc) ? 1 : 0;
      if (KEEP) rrow = a.res_idx[idx] + j;
    }
    s_idx[tid] = idx;
    s_loc[tid] = loc;
    s_flip[tid] = flip;
    s_res_row[tid] = rrow;
  }
  __syncthreads();

  // joined-chunk builder role: 8 threads per pair, 2 words (16 B) each
  const int b_pair = tid >> 3, b_sub = tid & 7;
  const uint32_t b_idx = s_idx[b_pair], b_loc = s_loc[b_pair];
  const bool b_valid = b_idx != 0xffffffffu;
  const bool b_flip = s_flip[b_pair] != 0;
  const uint64_t* b_p0 = a.p0 + (size_t)(b_valid ? b_idx : 0) * row_words;
  const uint64_t* b_p1 = a.p1 + (size_t)(b_valid ? b_loc : 0) * row_words;
  uint64_t* b_res = KEEP ? a.pres + (size_t)s_res_row[b_pair] * row_words : nullptr;
  const bool b_write = KEEP && b_valid && perm_tile == 0;
  unsigned cnt_in[M], cnt_out[M];
#pragma unroll
  for (int h = 0; h < M; h++) cnt_in[h] = cnt_out[h] = 0;

  unsigned acc[M][RP][RI];
#pragma unroll
  for (int h = 0; h < M; h++)
#pragma unroll
    for (int p = 0; p < RP; p++)
#pragma unroll
      for (int i = 0; i < RI; i++) acc[h][p][i] = 0;

  const int n_chunks = (Wp + KC - 1) / KC;
  const uint64_t* pm_tile = a.pm + (size_t)perm_tile * TI;

  auto stage_masks = [&](int chunk, int st) {
    // KC rows x TI*8 bytes = KC x 64 16-byte pieces
    const int k0 = chunk * KC;
    for (int i = tid; i < KC * (TI / 2); i += THREADS) {
      const int kr = i / (TI / 2), c = i % (TI / 2);
      if (k0 + kr < Wp) cp_async16(&s_mask[st][kr][c * 2], pm_tile + (size_t)(k0 + kr) * a.Ip + c * 2);
    }
    cp_async_commit();
  };

  stage_masks(0, 0);

  for (int chunk = 0; chunk < n_chunks; chunk++) {
    const int st = chunk & 1;
    const int k0 = chunk * KC;
    if (chunk + 1 < n_chunks) stage_masks(chunk + 1, st ^ 1);

    // ---- build joined chunk ----
    {
      const int k = k0 + b_sub * 2;
      ulonglong2 j[M];
#pragma unroll
      for (int h = 0; h < M; h++) j[h] = make_ulonglong2(0ull, 0ull);
      if (b_valid && k < Wp) {
        if (M == 1) {
          const ulonglong2 x = *reinterpret_cast<const ulonglong2*>(b_p0 + k);
          const ulonglong2 y = *reinterpret_cast<const ulonglong2*>(b_p1 + k);
          j[0] = make_ulonglong2(x.x | y.x, x.y | y.y);
        } else {
          // src/methods.h:137-145: flip -> downstream pos half is its first half, else its second
          const ulonglong2 x0 = *reinterpret_cast<const ulonglong2*>(b_p0 + k);
          const ulonglong2 x1 = *reinterpret_cast<const ulonglong2*>(b_p0 + Wp + k);
          const ulonglong2 y0 = *reinterpret_cast<const ulonglong2*>(b_p1 + (b_flip ? 0 : Wp) + k);
          const ulonglong2 y1 = *reinterpret_cast<const ulonglong2*>(b_p1 + (b_flip ? Wp : 0) + k);
          j[0] = make_ulonglong2(x0.x | y0.x, x0.y | y0.y);
          j[M - 1] = make_ulonglong2(x1.x | y1.x, x1.y | y1.y);
        }
        const uint64_t cm0 = case_mask_word(k, a.n_cases), cm1 = case_mask_word(k + 1, a.n_cases);
#pragma unroll
        for (int h = 0; h < M; h++) {
          cnt_in[h] += __popcll(j[h].x & cm0) + __popcll(j[h].y & cm1);
          cnt_out[h] += __popcll(j[h].x & ~cm0) + __popcll(j[h].y & ~cm1);
          if (b_write) *reinterpret_cast<ulonglong2*>(b_res + h * Wp + k) = j[h];
        }
      }
#pragma unroll
      for (int h = 0; h < M; h++) *reinterpret_cast<ulonglong2*>(&s_join[h][b_pair][b_sub * 2]) = j[h];
    }

    if (chunk + 1 < n_chunks) cp_async_wait<1>(); else cp_async_wait<0>();
    __syncthreads();

    // ---- AND + POPC over the chunk ----
    const int kmax = min(KC, Wp - k0);  // even
    for (int k = 0; k < kmax; k += 2) {
      ulonglong2 jw[M][RP];
#pragma unroll
      for (int h = 0; h < M; h++)
#pragma unroll
        for (int p = 0; p < RP; p++) jw[h][p] = *reinterpret_cast<const ulonglong2*>(&s_join[h][warp * RP + p][k]);
#pragma unroll
      for (int kk = 0; kk < 2; kk++) {
        const ulonglong2 m01 = *reinterpret_cast<const ulonglong2*>(&s_mask[st][k + kk][lane * 2]);
        const ulonglong2 m23 = *reinterpret_cast<const ulonglong2*>(&s_mask[st][k + kk][64 + lane * 2]);
#pragma unroll
        for (int h = 0; h < M; h++)
#pragma unroll
          for (int p = 0; p < RP; p++) {
            const uint64_t jv = kk ? jw[h][p].y : jw[h][p].x;
            acc[h][p][0] += __popcll(jv & m01.x);
            acc[h][p][1] += __popcll(jv & m01.y);
            acc[h][p][2] += __popcll(jv & m23.x);
            acc[h][p][3] += __popcll(jv & m23.y);
          }
      }
    }
    __syncthreads();
  }

  // ---- true counts per pair: reduce over the 8 builder threads of the pair ----
#pragma unroll
  for (int h = 0; h < M; h++) {
    unsigned x = cnt_in[h], y = cnt_out[h];
    x += __shfl_xor_sync(0xffffffffu, x, 1); y += __shfl_xor_sync(0xffffffffu, y, 1);
    x += __shfl_xor_sync(0xffffffffu, x, 2); y += __shfl_xor_sync(0xffffffffu, y, 2);
    x += __shfl_xor_sync(0xffffffffu, x, 4); y += __shfl_xor_sync(0xffffffffu, y, 4);
    if (b_sub == 0) { s_cnt[b_pair][h][0] = x; s_cnt[b_pair][h][1] = y; }
  }
  __syncthreads();

  // ---- permutation look-ups + per-perm max ----
  // perms of this thread: lane*2, lane*2+1, 64+lane*2, 64+lane*2+1 (within the tile)
  float* s_best = reinterpret_cast<float*>(&s_mask[0][0][0]);  // [WARPS][TI] floats, mask buffers are free now
  if (M == 1) {
    float best[RI];
#pragma unroll
    for (int i = 0; i < RI; i++) best[i] = 0.0f;
#pragma unroll
    for (int p = 0; p < RP; p++) {
      const int pl = warp * RP + p;
      if (s_idx[pl] == 0xffffffffu) continue;
      const unsigned total = s_cnt[pl][0][0] + s_cnt[pl][0][1];
      const float* row = a.diagF + diag_base(total);
#pragma unroll
      for (int i = 0; i < RI; i++) best[i] = fmaxf(best[i], __ldg(row + acc[0][p][i]));
    }
    s_best[warp * TI + lane * 2] = best[0];
    s_best[warp * TI + lane * 2 + 1] = best[1];
    s_best[warp * TI + 64 + lane * 2] = best[2];
    s_best[warp * TI + 64 + lane * 2 + 1] = best[3];
  } else {
    double best[RI];
#pragma unroll
    for (int i = 0; i < RI; i++) best[i] = 0.0;
#pragma unroll
    for (int p = 0; p < RP; p++) {
      const int pl = warp * RP + p;
      if (s_idx[pl] == 0xffffffffu) continue;
      // src/methods.h:220-230: vtmax[pcp][total_pos - pcp] + vtmax[total_neg - pnp][pnp]
      const unsigned tpos = s_cnt[pl][0][0] + s_cnt[pl][0][1];
      const unsigned tneg = s_cnt[pl][M - 1][0] + s_cnt[pl][M - 1][1];
      const double* rowp = a.diagDM + diag_base(tpos);
      const double* rown = a.diagDM + diag_base(tneg);
#pragma unroll
      for (int i = 0; i < RI; i++) {
        const double v = __ldg(rowp + acc[0][p][i]) + __ldg(rown + (tneg - acc[M - 1][p][i]));
        best[i] = fmax(best[i], v);
      }
    }
    s_best[warp * TI + lane * 2] = __double2float_rn(best[0]);
    s_best[warp * TI + lane * 2 + 1] = __double2float_rn(best[1]);
    s_best[warp * TI + 64 + lane * 2] = __double2float_rn(best[2]);
    s_best[warp * TI + 64 + lane * 2 + 1] = __double2float_rn(best[3]);
  }
  __syncthreads();
  if (tid < TI) {
    float v = s_best[tid];
#pragma unroll
    for (int w = 1; w < WARPS; w++) v = fmaxf(v, s_best[w * TI + tid]);
    if (v > 0.0f) atomicMax(a.perm_max + perm_tile * TI + tid, __float_as_int(v));
  }

  // ---- true score of each pair -> top-K candidates (one CTA per pair tile does it) ----
  if (perm_tile == 0 && tid < TP && s_idx[tid] != 0xffffffffu) {
    double score;
    int cases, ctrls;
    unsigned tmax;
    if (M == 1) {
      // src/methods.h:90: vt[cases][ctrls]
      cases = (int)s_cnt[tid][0][0];
      ctrls = (int)s_cnt[tid][0][1];
      tmax = (unsigned)(cases + ctrls);
      score = a.diagD[diag_base(tmax) + cases];
    } else {
      // src/methods.h:253-257: vt[case_pos][ctrl_neg] + vt[case_neg][ctrl_pos]
      const unsigned case_pos = s_cnt[tid][0][0], ctrl_neg = s_cnt[tid][0][1];
      const unsigned ctrl_pos = s_cnt[tid][M - 1][0], case_neg = s_cnt[tid][M - 1][1];
      score = a.diagD[diag_base(case_pos + ctrl_neg) + case_pos] + a.diagD[diag_base(case_neg + ctrl_pos) + case_neg];
      cases = (int)(case_pos + case_neg);
      ctrls = (int)(ctrl_pos + ctrl_neg);
      tmax = max(case_pos + ctrl_neg, case_neg + ctrl_pos);
    }
    if (KEEP) atomicMax(a.max_total, tmax);
    if (score == score) {
      const unsigned long long key = score_key(score);
      if (key > a.thr_key) {
        const unsigned slot = atomicAdd(a.cand_count, 1u);
        if (slot < a.cand_cap) {
          Cand c;
          c.key = key; c.idx = s_idx[tid]; c.loc = s_loc[tid]; c.cases = cases; c.ctrls = ctrls;
          a.cand[slot] = c;
        }
      }
    }
  }
}

}  // namespace gcre
