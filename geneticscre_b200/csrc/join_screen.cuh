// Range-bound screening of score-only sparse joins (sm_100a).
//
// A score-only join needs, per permutation, only the MAXIMUM of the pair scores.  Almost every pair is nowhere near it: the
// running maximum over millions of pairs sits at -log p ~ 15, a typical pair scores ~1.  This kernel decides per pair, from
// integer RANGES instead of the 1,024 exact counts, that no permutation of the pair can beat the maxima already established,
// and hands the few pairs it cannot rule out to the exact kernel (join_sparse.cuh).  Results stay bit-identical: a pair is
// dropped only when an upper bound of all its permutation scores is <= a lower bound of the current maxima.
//
//   count[r] = base[r] + delta[r]            base: the upstream row (counts emitted by the join that made it),
//                                            delta: the partner's carriers not in the upstream row (bit planes, join_sparse.cuh)
//   lane l owns 32 permutations:  cmin = min_r base + min_r delta <= count[r] <= max_r base + max_r delta = cmax
//     min_r / max_r base : stored with the emitted counts (one u32 per lane, SparseView::prange)
//     min_r / max_r delta: read off the bit planes by a bit-sliced max / min (4 instructions per plane) - the planes are never
//                          expanded into counters
//   score[r] = T[total][count[r]] <= max(EL[total][cmin], ER[total][cmax])
//     EL / ER: envelopes of every table row, built once per exec: EL[c] = max T[c..m], ER[c] = max T[m..c] for a split point m
//     (any m is valid; the expected count is tight for the U-shaped rows of -log p tables).  Method 2 adds the two halves'
//     bounds with round-up float arithmetic over round-up envelopes of the f64 table, so the bound dominates the f64 sum and
//     its f32 rounding.
//   thr[l] = min over the lane's real permutations of the maxima established by a SEED pass (the exact kernel over every
//     64th unit); maxima only grow, so bound <= thr[l] proves the pair cannot change any of the lane's maxima.
//
// Launch sequence of a screened join (all stream ordered, no host round trip): seed list -> exact kernel on the seed ->
// thresholds -> this kernel (true scores, top-K candidates, retry list) -> exact kernel on the retry list.
// Permutation slots beyond the requested count are filled with copies of real permutations in the patient-major masks
// (masks_to_patient_major_kernel), so they never widen a range; their maxima are never read.
#pragma once
#include "join_sparse.cuh"

namespace gcre {

namespace screen {
constexpr int THREADS = 128;
constexpr int WARPS = THREADS / 32;
constexpr int MIN_BLOCKS = 8;      // 64 registers
constexpr int SEED_STRIDE = 64;    // every 64th unit is scored exactly before the screening pass
constexpr unsigned long long MIN_UNITS = 4096;  // smaller joins run the exact kernel directly
}  // namespace screen

// min / max over the 32 permutations of a lane of the counts held in bit planes (plane j = bit j of every count);
// nbits: planes that can be populated (warp-uniform)
__device__ __forceinline__ void plane_range(const uint32_t (&pl)[8], int nbits, uint32_t& mn, uint32_t& mx) {
  uint32_t cmx = 0xffffffffu, cmn = 0xffffffffu;
  mn = 0;
  mx = 0;
#pragma unroll
  for (int j = 7; j >= 0; j--) {
    if (j < nbits) {
      const uint32_t t = cmx & pl[j];   // candidates for the maximum that have bit j set
      if (t) { cmx = t; mx |= 1u << j; }
      const uint32_t u = cmn & ~pl[j];  // candidates for the minimum that have bit j clear
      if (u) cmn = u; else mn |= 1u << j;
    }
  }
}

// ---- envelopes of the permutation look-up tables ------------------------------------------------------------------------
// One CTA per anti-diagonal row t (entries c = 0..t at diag_base(t)).  SRC = float: method 1 (F, already the f32 the exact
// kernel looks up); SRC = double: method 2 (DM; envelopes rounded UP to float).  NaN entries never win a maximum in the exact
// kernel (fmaxf / the reference's `>`), so they are skipped here too.
template <typename SRC>
__global__ void __launch_bounds__(128) build_env_kernel(const SRC* __restrict__ T, unsigned t_cap, int n_cases, int n, float2* __restrict__ env) {
  __shared__ float s_part[128];
  const unsigned t = blockIdx.x;
  if (t > t_cap) return;
  const size_t base = diag_base(t);
  const unsigned m = min(t, (unsigned)(((unsigned long long)t * (unsigned)n_cases + (unsigned)n / 2) / (unsigned)n));  // split point
  auto val = [&](unsigned c) -> float {
    const SRC v = T[base + c];
    if (v != v) return -INFINITY;
    if (sizeof(SRC) == 8) return __double2float_ru((double)v);
    return (float)v;
  };
  // left: EL[c] = max T[c..m] (c <= m), a running maximum from m down to 0; right: ER[c] = max T[m..c] (c >= m)
  for (int side = 0; side < 2; side++) {
    const unsigned len = side == 0 ? m + 1 : t - m + 1;            // positions k = 0..len-1 counted from m outwards
    const unsigned chunk = (len + 127) / 128;
    const unsigned k0 = min(len, threadIdx.x * chunk), k1 = min(len, k0 + chunk);
    float acc = -INFINITY;
    for (unsigned k = k0; k < k1; k++) acc = fmaxf(acc, val(side == 0 ? m - k : m + k));
    s_part[threadIdx.x] = acc;
    __syncthreads();
    float run = -INFINITY;
    for (unsigned q = 0; q < threadIdx.x; q++) run = fmaxf(run, s_part[q]);
    for (unsigned k = k0; k < k1; k++) {
      const unsigned c = side == 0 ? m - k : m + k;
      run = fmaxf(run, val(c));
      if (side == 0) env[base + c].x = run; else env[base + c].y = run;
    }
    __syncthreads();
  }
  // outside its side an envelope repeats its value at the split point: max(EL[cmin], ER[cmax]) needs no clamping
  const float at_m = val(m);
  for (unsigned c = threadIdx.x; c <= t; c += 128) {
    if (c > m) env[base + c].x = at_m;
    if (c < m) env[base + c].y = at_m;
  }
}

// thr[pb * 32 + lane] = min over the REAL permutations of lane `lane` of block `pb` of the current maxima (+inf if none)
__global__ void screen_thresholds_kernel(const int* __restrict__ perm_max, int iters, int n_words, float* __restrict__ thr) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_words) return;
  float m = INFINITY;
  for (int b = 0; b < 32; b++) {
    const int r = w * 32 + b;
    if (r < iters) m = fminf(m, __int_as_float(perm_max[r]));
  }
  thr[w] = m;
}

// every `stride`-th unit of the launch, all permutation blocks, all partners
__global__ void screen_seed_list_kernel(unsigned long long unit_begin, unsigned long long n_units, int stride, int n_perm_blocks,
                                        RetryEntry* __restrict__ out, unsigned* __restrict__ count) {
  const unsigned long long n_seed = (n_units + stride - 1) / stride;
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_seed * n_perm_blocks) return;
  RetryEntry e;
  e.unit = (uint32_t)(unit_begin + (i % n_seed) * stride);
  e.pb = (uint32_t)(i / n_seed);
  e.mask = ~0ull;
  out[i] = e;
  if (i == 0) *count = (unsigned)(n_seed * n_perm_blocks);
}

struct ScreenParams {
  const uint32_t* prange0;   // [upstream item][perm block][32] min | max << 16 of the base counts
  const float2* env;         // envelopes (x = EL, y = ER), anti-diagonal-major like the look-up tables
  const float* thr;          // [n_perm_blocks * 32]
  RetryEntry* retry;         // capacity n_units * n_perm_blocks
  unsigned* retry_count;
};

template <int M, typename CT>
__global__ void __launch_bounds__(screen::THREADS, screen::MIN_BLOCKS) join_screen_kernel(const JoinParams a, const SparseParams s, const ScreenParams z) {
  using namespace screen;
  constexpr int PB = sparse::PB, FLUSH_AT = sparse::FLUSH_AT, QCAP = sparse::QCAP;
  const CT* car1 = static_cast<const CT*>(s.car1);
  __shared__ __align__(16) uint32_t s_queue[WARPS][QCAP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Wp = a.Wp, Iw = a.Iw;
  const int row_words = Wp * M;
  const unsigned long long n_work = s.n_units * (unsigned long long)s.n_perm_blocks;
  const unsigned lt_mask = (1u << lane) - 1u;
  uint32_t* queue = s_queue[warp];

  while (true) {
    unsigned long long g = 0;
    if (lane == 0) g = atomicAdd(s.work_counter, 1ull);
    g = __shfl_sync(0xffffffffu, g, 0);
    if (g >= n_work) break;
    const int pb = (int)(g / s.n_units);
    const unsigned long long unit = s.unit_begin + (g % s.n_units);
    const bool first_pb = (pb == 0);
    const uint32_t idx = s.unit_idx[unit];
    const uint32_t sub = (uint32_t)(unit - s.unit_prefix[idx]);
    const uint32_t cnt_idx = (uint32_t)a.count[idx];
    const uint32_t j0 = sub * PB, j1 = min(cnt_idx, j0 + PB);
    const uint32_t* pt_lane = a.pt + pb * 32 + lane;
    const uint64_t* p0row = a.p0 + (size_t)idx * row_words;
    const uint32_t loc0 = a.location[idx];
    const float thr = z.thr[pb * 32 + lane];

    uint32_t t0[M], nc0[M], brange[M];
#pragma unroll
    for (int h = 0; h < M; h++) {
      const size_t item = (size_t)idx * M + h;
      t0[h] = s.len0[item];
      nc0[h] = s.ncase0[item];
      brange[h] = __ldg(z.prange0 + (item * s.n_perm_blocks + pb) * 32 + lane);
    }

    uint32_t pl[8];
#pragma unroll
    for (int j = 0; j < 8; j++) pl[j] = 0;
    int inbatch = 0, inreal = 0;
    unsigned long long retry_mask = 0;  // lane 0's copy is the one that is published
    bool base_done = false;

    for (uint32_t j = j0; j < j1; j++) {
      const uint32_t loc = loc0 + j;
      bool flip = true;
      if (M == 2) flip = need_flip(a.path_length, a.signs, idx, loc);
      uint32_t nd[M], ncn[M], dmn[M], dmx[M];
#pragma unroll
      for (int h = 0; h < M; h++) nd[h] = ncn[h] = dmn[h] = dmx[h] = 0;
#pragma unroll 1
      for (int h = 0; h < M; h++) {
        const int hh = (M == 1) ? 0 : (flip ? h : 1 - h);
        const size_t item = (size_t)loc * M + hh;
        const CT* lst1 = car1 + s.off1[item];
        const uint32_t len = s.len1[item];
        const uint64_t* p0h = p0row + h * Wp;
        uint32_t ndh = 0, ncnh = 0, mnS = 0, mxS = 0;
        uint32_t qn = 0;
#pragma unroll 1
        for (uint32_t i0 = 0; i0 < len; i0 += 32) {
          const uint32_t i = i0 + lane;
          const bool valid = i < len;
          const uint32_t c = valid ? (uint32_t)lst1[i] : 0u;
          const uint32_t w0 = valid ? __ldg(reinterpret_cast<const uint32_t*>(p0h) + (c >> 5)) : 0u;
          const bool keep = valid && !((w0 >> (c & 31)) & 1u);
          const unsigned km = __ballot_sync(0xffffffffu, keep);
          ncnh += __popc(__ballot_sync(0xffffffffu, keep && (int)c < a.n_cases));
          if (keep) queue[qn + __popc(km & lt_mask)] = c * (uint32_t)Iw;
          qn += __popc(km);
          const bool last_chunk = i0 + 32 >= len;
          if (qn >= 64 || (last_chunk && (qn > 0 || inbatch > 0))) {
            const uint32_t take = last_chunk ? max((qn + 7u) & ~7u, 8u) : 64u;
            if (last_chunk && lane < take - qn) queue[qn + lane] = (uint32_t)s.n * (uint32_t)Iw;
            __syncwarp();
#pragma unroll 1
            for (uint32_t q0 = 0; q0 < take; q0 += 8) {
              const uint4 lo = *reinterpret_cast<const uint4*>(queue + q0), hi = *reinterpret_cast<const uint4*>(queue + q0 + 4);
              uint32_t x[8];
              x[0] = __ldg(pt_lane + lo.x);
              x[1] = __ldg(pt_lane + lo.y);
              x[2] = __ldg(pt_lane + lo.z);
              x[3] = __ldg(pt_lane + lo.w);
              x[4] = __ldg(pt_lane + hi.x);
              x[5] = __ldg(pt_lane + hi.y);
              x[6] = __ldg(pt_lane + hi.z);
              x[7] = __ldg(pt_lane + hi.w);
              hs8(pl, x);
              inbatch += 8;
              inreal += (int)min(8u, qn > q0 ? qn - q0 : 0u);
              if (inbatch > FLUSH_AT || (last_chunk && q0 + 8 >= take)) {
                // fold the planes into the running range of the delta: the sum of per-batch minima / maxima bounds the total
                uint32_t mn, mx;
                plane_range(pl, inreal ? bits_for(inreal) : 0, mn, mx);
                mnS += mn;
                mxS += mx;
#pragma unroll
                for (int q = 0; q < 8; q++) pl[q] = 0;
                inbatch = inreal = 0;
              }
            }
            const uint32_t rem = last_chunk ? 0u : qn - 64u;
            const uint32_t keepv = (lane < rem) ? queue[64 + lane] : 0u;
            __syncwarp();
            if (lane < rem) queue[lane] = keepv;
            ndh += qn - rem;
            qn = rem;
          }
        }
        if (M == 1 || h == 0) { nd[0] = ndh; ncn[0] = ncnh; dmn[0] = mnS; dmx[0] = mxS; }
        else { nd[M - 1] = ndh; ncn[M - 1] = ncnh; dmn[M - 1] = mnS; dmx[M - 1] = mxS; }
      }

      bool empty = nd[0] == 0;
      if (M == 2) empty = empty && nd[M - 1] == 0;
      if (!empty || !base_done) {
        if (empty) base_done = true;
        float bound;
        if (M == 1) {
          const unsigned total = t0[0] + nd[0];
          const float2* row = z.env + diag_base(total);
          const uint32_t cmin = (brange[0] & 0xffffu) + dmn[0], cmax = (brange[0] >> 16) + dmx[0];
          bound = fmaxf(__ldg(&row[cmin].x), __ldg(&row[cmax].y));
        } else {
          // src/methods.h:223-227: vtmax[pcp][total_pos - pcp] + vtmax[total_neg - pnp][pnp]: the second look-up runs down its row
          const unsigned tp = t0[0] + nd[0], tn = t0[M - 1] + nd[M - 1];
          const float2* rowp = z.env + diag_base(tp);
          const float2* rown = z.env + diag_base(tn);
          const uint32_t cpmin = (brange[0] & 0xffffu) + dmn[0], cpmax = (brange[0] >> 16) + dmx[0];
          const uint32_t cnmin = (brange[M - 1] & 0xffffu) + dmn[M - 1], cnmax = (brange[M - 1] >> 16) + dmx[M - 1];
          const float bp = fmaxf(__ldg(&rowp[cpmin].x), __ldg(&rowp[cpmax].y));
          const float bn = fmaxf(__ldg(&rown[tn - cnmax].x), __ldg(&rown[tn - cnmin].y));
          bound = __fadd_ru(bp, bn);
        }
        if (__any_sync(0xffffffffu, bound > thr)) retry_mask |= 1ull << (j - j0);
      }

      if (first_pb && lane == 0) {
        // ---- true score -> top-K candidate (src/methods.h:90-94, 253-264), as in join_sparse_kernel ----
        double score;
        int cases, ctrls;
        if (M == 1) {
          cases = (int)(nc0[0] + ncn[0]);
          const unsigned tmax = t0[0] + nd[0];
          ctrls = (int)tmax - cases;
          score = a.diagD[diag_base(tmax) + cases];
        } else {
          const unsigned tp = t0[0] + nd[0], tn = t0[M - 1] + nd[M - 1];
          const unsigned case_pos = nc0[0] + ncn[0], ctrl_neg = tp - case_pos;
          const unsigned ctrl_pos = nc0[M - 1] + ncn[M - 1], case_neg = tn - ctrl_pos;
          score = a.diagD[diag_base(tp) + case_pos] + a.diagD[diag_base(tn) + case_neg];
          cases = (int)(case_pos + case_neg);
          ctrls = (int)(ctrl_pos + ctrl_neg);
        }
        if (score == score) {
          const unsigned long long key = score_key(score);
          const unsigned long long dyn = a.n_slots ? __ldcg(a.dyn_thr) : 0ull;
          if (key > a.thr_key && key >= dyn) {
            const unsigned slot = atomicAdd(a.cand_count, 1u);
            if (slot < a.cand_cap) {
              Cand cd;
              cd.key = key; cd.idx = idx; cd.loc = loc; cd.cases = cases; cd.ctrls = ctrls;
              a.cand[slot] = cd;
            }
            if (a.n_slots) {
              const unsigned bucket = ((idx * 0x9E3779B1u) ^ (loc * 0x85EBCA6Bu)) >> 8;
              if (atomicMax(a.slots + bucket % (unsigned)a.n_slots, key) < key) {
                unsigned long long m = ~0ull;
                for (int t = 0; t < a.n_slots; t++) m = min(m, __ldcg(a.slots + t));
                if (m > dyn) atomicMax(a.dyn_thr, m);
              }
            }
          }
        }
      }
    }
    // partners that added nothing share the base evaluation: if that one was flagged, all of them are (cheap: the exact
    // kernel evaluates the base once per unit as well)
    if (lane == 0 && retry_mask) {
      const unsigned slot = atomicAdd(z.retry_count, 1u);
      RetryEntry e;
      e.unit = (uint32_t)unit;
      e.pb = (uint32_t)pb;
      e.mask = retry_mask;
      z.retry[slot] = e;
    }
    __syncwarp();
  }
}

template <int M>
static inline cudaError_t launch_join_screen(cudaStream_t stream, const JoinParams& jp, const SparseParams& sp, const ScreenParams& zp, int sm_count) {
  const unsigned long long n_work = sp.n_units * (unsigned long long)sp.n_perm_blocks;
  if (n_work == 0) return cudaSuccess;
  const unsigned long long want = (n_work + screen::WARPS - 1) / screen::WARPS;
  const unsigned grid = (unsigned)std::min<unsigned long long>(want, (unsigned long long)sm_count * screen::MIN_BLOCKS);
  if (sparse_wide(sp.n)) join_screen_kernel<M, uint32_t><<<grid, screen::THREADS, 0, stream>>>(jp, sp, zp);
  else join_screen_kernel<M, uint16_t><<<grid, screen::THREADS, 0, stream>>>(jp, sp, zp);
  return cudaGetLastError();
}

}  // namespace gcre
