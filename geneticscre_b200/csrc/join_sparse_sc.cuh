// Sparse join for FEW permutations (<= 512, e.g. GWASPA's default of 100): the carriers of one pair are split across the
// lanes of the warp.
//
// join_sparse_kernel gives every lane one 32-permutation word of the patient-major masks; with <= 128 permutations only 4
// of the 32 lanes have a word and 28 idle through the gathers, the flush and the look-ups (8 of 32 busy up to 256, 16 up to
// 512).  Here the warp is G = 32 / NW sub-groups x NW words (NW = 4, 8, 16; w = lane % NW, s = lane / NW): sub-group s walks
// every G-th group of eight carriers of the SAME pair, so a batch of 8 G carriers costs each lane eight gathers; the G partial
// count vectors of a word are then combined with a reduce-scatter over shuffles (log2 G rounds), which leaves lane (w, s)
// with 16 / G of the 16 packed count registers - for NW = 4 the counts of the four permutations w*32 + {s, s+8, s+16, s+24} -
// the only ones it looks up and keeps a running maximum for.  Same integers as every other
// kernel, so results are bit-identical.  (Splitting the warp over different PARTNERS instead was tried first and lost to
// the heavy-tailed list lengths; splitting one pair's carriers has no such imbalance.)
//
// Everything else - units, filter against the dense upstream row, true scores, candidates with the self-tightening
// threshold, kept rows - is as in join_sparse.cuh.  No count tables are emitted or consumed here: the base of a unit is
// walked the same split way and costs an eighth of what it does there.
#pragma once
#include "join_sparse.cuh"

namespace gcre {

namespace sparse_sc {
constexpr int THREADS = 128;
constexpr int WARPS = THREADS / 32;
constexpr int QCAP = 128;       // the batch being drained (<= 64 slots) + the < 32 entries that can wait behind it, in whole batches
constexpr int MAX_PERMS = 512;  // 16 words of 32
constexpr int min_blocks(int nw) { return nw == 16 ? 6 : 8; }
}  // namespace sparse_sc

// position of the t-th queued carrier: inside its batch of B = 8 G the entries of sub-group s = t % G are contiguous (two
// 16-byte shared loads per lane and batch)
template <int G>
__device__ __forceinline__ uint32_t sc_slot(uint32_t t) {
  constexpr uint32_t B = 8 * G;
  return (t / B) * B + ((t % G) << 3) + ((t % B) / G);
}

// 16 packed count registers per lane (partial sums over the lane's share of the carriers) -> the R = NW / 2 registers
// R*s .. R*s + R-1 summed over the G sub-groups of the lane's word (s = lane / NW)
template <int NW>
__device__ __forceinline__ void sc_reduce_scatter(const uint32_t (&c16)[16], int lane, uint32_t (&out)[NW / 2]) {
  const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
  uint32_t r8[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const uint32_t send = b4 ? c16[k] : c16[k + 8], keep = b4 ? c16[k + 8] : c16[k];
    r8[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  if constexpr (NW == 16) {
#pragma unroll
    for (int k = 0; k < 8; k++) out[k] = r8[k];
  } else {
    uint32_t r4[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t send = b3 ? r8[k] : r8[k + 4], keep = b3 ? r8[k + 4] : r8[k];
      r4[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    if constexpr (NW == 8) {
#pragma unroll
      for (int k = 0; k < 4; k++) out[k] = r4[k];
    } else {
#pragma unroll
      for (int k = 0; k < 2; k++) {
        const uint32_t send = b2 ? r4[k] : r4[k + 2], keep = b2 ? r4[k + 2] : r4[k];
        out[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
    }
  }
}

template <int M, bool KEEP, typename CT, int NW>
__global__ void __launch_bounds__(sparse_sc::THREADS, sparse_sc::min_blocks(NW)) join_sparse_sc_kernel(const JoinParams a, const SparseParams s) {
  using namespace sparse_sc;
  constexpr int PB = sparse::PB, FLUSH_AT = sparse::FLUSH_AT;
  constexpr int G = 32 / NW;          // sub-groups sharing one pair's carriers
  constexpr uint32_t B = 8 * G;       // carriers per batch
  constexpr int R = NW / 2;           // packed count registers (2 permutations each) a lane keeps after the reduce-scatter
  constexpr int NCOPY = (B == 64) ? 2 : 1;  // 32-slot pieces that can hold the < 32 entries left behind a drained batch
  const CT* car0 = static_cast<const CT*>(s.car0);
  const CT* car1 = static_cast<const CT*>(s.car1);
  __shared__ __align__(16) uint32_t s_queue[WARPS][QCAP];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int w = lane % NW, sub = lane / NW;
  const int Wp = a.Wp, Iw = a.Iw;
  const int row_words = Wp * M;
  const unsigned lt_mask = (1u << lane) - 1u;
  uint32_t* queue = s_queue[warp];
  const uint32_t* pt_lane = a.pt + w;
  const uint32_t zero_row = (uint32_t)s.n * (uint32_t)Iw;

  // permutations of this lane: (register i = R*sub + k, half hf)  <->  bit ((i & 1)*2 + hf)*8 + (i >> 1) of word w
  float best[2 * R];
#pragma unroll
  for (int j = 0; j < 2 * R; j++) best[j] = 0.0f;

  while (true) {
    unsigned long long g = 0;
    if (lane == 0) g = atomicAdd(s.work_counter, 1ull);
    g = __shfl_sync(0xffffffffu, g, 0);
    if (g >= s.n_units) break;
    const unsigned long long unit = s.unit_begin + g;
    const uint32_t idx = s.unit_idx[unit];
    const uint32_t subu = (uint32_t)(unit - s.unit_prefix[idx]);
    const uint32_t cnt_idx = (uint32_t)a.count[idx];
    const uint32_t j0 = subu * PB, j1 = min(cnt_idx, j0 + PB);
    const uint64_t* p0row = a.p0 + (size_t)idx * row_words;
    const uint32_t loc0 = a.location[idx];

    uint32_t pl[8];
#pragma unroll
    for (int j = 0; j < 8; j++) pl[j] = 0;
    int inbatch = 0, inreal = 0;

    auto acc8 = [&](uint32_t (&c16)[16], const uint32_t (&x)[8], int real, bool last) {
      hs8(pl, x);
      inbatch += 8;
      inreal += real;
      if (inbatch > FLUSH_AT || last) {
        flush_planes(c16, pl, bits_for(inreal));
        inbatch = inreal = 0;
      }
    };

    // ---- base: the upstream row's own carriers, 64 per step, sub-group `sub` takes the sub-th group of eight ----
    uint32_t t0[M], nc0[M], base[M][R];
#pragma unroll
    for (int h = 0; h < M; h++) {
      t0[h] = nc0[h] = 0;
#pragma unroll
      for (int k = 0; k < R; k++) base[h][k] = 0;
    }
#pragma unroll 1
    for (int h = 0; h < M; h++) {
      uint32_t acc[16];
#pragma unroll
      for (int i = 0; i < 16; i++) acc[i] = 0;
      const size_t item = (size_t)idx * M + h;
      const CT* lst0 = car0 + s.off0[item];
      const uint32_t plen = (uint32_t)(s.off0[item + 1] - s.off0[item]);
      const uint32_t t0h = s.len0[item], nc0h = s.ncase0[item];
#pragma unroll 1
      for (uint32_t i = 0; i < plen; i += B) {
        uint32_t x[8];
        if (i + sub * 8 < plen) {
          uint32_t c[8];
          load8(lst0 + i + sub * 8, c);
#pragma unroll
          for (int q = 0; q < 8; q++) x[q] = __ldg(word_ptr(pt_lane, c[q] * (uint32_t)Iw));
        } else {
#pragma unroll
          for (int q = 0; q < 8; q++) x[q] = 0u;
        }
        acc8(acc, x, 8, i + B >= plen);
      }
      uint32_t rr[R];
      sc_reduce_scatter<NW>(acc, lane, rr);
      if (M == 1 || h == 0) {
        t0[0] = t0h; nc0[0] = nc0h;
#pragma unroll
        for (int k = 0; k < R; k++) base[0][k] = rr[k];
      } else {
        t0[M - 1] = t0h; nc0[M - 1] = nc0h;
#pragma unroll
        for (int k = 0; k < R; k++) base[M - 1][k] = rr[k];
      }
    }

    bool base_done = false;
    // ---- partners ----
    for (uint32_t j = j0; j < j1; j++) {
      const uint32_t loc = loc0 + j;
      bool flip = true;
      if (M == 2) flip = need_flip(a.path_length, a.signs, idx, loc);
      uint32_t nd[M], ncn[M], cnt[M][R];
#pragma unroll
      for (int h = 0; h < M; h++) {
        nd[h] = ncn[h] = 0;
#pragma unroll
        for (int k = 0; k < R; k++) cnt[h][k] = 0;
      }
#pragma unroll 1
      for (int h = 0; h < M; h++) {
        uint32_t c16[16];
#pragma unroll
        for (int i = 0; i < 16; i++) c16[i] = 0;
        const int hh = (M == 1) ? 0 : (flip ? h : 1 - h);
        const size_t item = (size_t)loc * M + hh;
        const CT* lst1 = car1 + s.off1[item];
        const uint32_t len = s.len1[item];
        const uint64_t* p0h = p0row + h * Wp;
        uint32_t ndh = 0, ncnh = 0;
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
          if (keep) queue[sc_slot<G>(qn + __popc(km & lt_mask))] = c * (uint32_t)Iw;
          qn += __popc(km);
          const bool last_chunk = i0 + 32 >= len;
          // drain whole batches of B; after the last chunk also the remainder (padded with the zero row), and at least once
          // when counts of an earlier batch are still in the planes
#pragma unroll 1
          while (qn >= B || (last_chunk && (qn > 0 || inbatch > 0))) {
            const uint32_t real = min(qn, B);
            if (real < B) {
              for (uint32_t t = real + lane; t < B; t += 32) queue[sc_slot<G>(t)] = zero_row;
            }
            __syncwarp();
            uint32_t x[8];
            const uint4 lo = *reinterpret_cast<const uint4*>(queue + sub * 8), hi = *reinterpret_cast<const uint4*>(queue + sub * 8 + 4);
            x[0] = __ldg(word_ptr(pt_lane, lo.x));
            x[1] = __ldg(word_ptr(pt_lane, lo.y));
            x[2] = __ldg(word_ptr(pt_lane, lo.z));
            x[3] = __ldg(word_ptr(pt_lane, lo.w));
            x[4] = __ldg(word_ptr(pt_lane, hi.x));
            x[5] = __ldg(word_ptr(pt_lane, hi.y));
            x[6] = __ldg(word_ptr(pt_lane, hi.z));
            x[7] = __ldg(word_ptr(pt_lane, hi.w));
            const uint32_t rem = qn - real;                      // < 32 entries waiting behind the drained batch
            acc8(c16, x, (int)((real + G - 1) / G), last_chunk && rem == 0);
            // they move to the front: the slot layout is per batch, so whole batches are copied as they are
            uint32_t mv[NCOPY];
#pragma unroll
            for (int k = 0; k < NCOPY; k++) mv[k] = queue[B + 32 * k + lane];
            __syncwarp();
            if (rem) {
#pragma unroll
              for (int k = 0; k < NCOPY; k++) queue[32 * k + lane] = mv[k];
            }
            __syncwarp();
            ndh += real;
            qn = rem;
          }
        }
        uint32_t rr[R];
        sc_reduce_scatter<NW>(c16, lane, rr);
        if (M == 1 || h == 0) {
          nd[0] = ndh; ncn[0] = ncnh;
#pragma unroll
          for (int k = 0; k < R; k++) cnt[0][k] = base[0][k] + rr[k];
        } else {
          nd[M - 1] = ndh; ncn[M - 1] = ncnh;
#pragma unroll
          for (int k = 0; k < R; k++) cnt[M - 1][k] = base[M - 1][k] + rr[k];
        }
      }

      bool empty = nd[0] == 0;
      if (M == 2) empty = empty && nd[M - 1] == 0;
      if (!empty || !base_done) {
        if (empty) base_done = true;
        if (M == 1) {
          const unsigned total = t0[0] + nd[0];
          const float* row = a.diagF + diag_base(total);
#pragma unroll
          for (int q = 0; q < 2 * R; q++) {
            const uint32_t v = cnt[0][q >> 1];
            const uint32_t c = (q & 1) ? (v >> 16) : (v & 0xffffu);
            best[q] = fmaxf(best[q], lookup_f32(row, c));
          }
        } else {
          const unsigned tp = t0[0] + nd[0], tn = t0[M - 1] + nd[M - 1];
          const double* rowp = a.diagDM + diag_base(tp);
          const double* rown = a.diagDM + diag_base(tn);  // indexed by tn - cn
#pragma unroll
          for (int q = 0; q < 2 * R; q++) {
            const uint32_t vp = cnt[0][q >> 1], vn = cnt[M - 1][q >> 1];
            const uint32_t cp = (q & 1) ? (vp >> 16) : (vp & 0xffffu);
            const uint32_t cn = (q & 1) ? (vn >> 16) : (vn & 0xffffu);
            // src/methods.h:223-227: vtmax[pcp][total_pos - pcp] + vtmax[total_neg - pnp][pnp]
            const double v = lookup_f64(rowp, cp) + lookup_f64(rown, tn - cn);
            best[q] = fmaxf(best[q], __double2float_rn(v));
          }
        }
      }

      // ---- kept joined row (src/join_base.cpp:246-249) ----
      if (KEEP) {
        const ulonglong2* u = reinterpret_cast<const ulonglong2*>(p0row);
        const ulonglong2* v = reinterpret_cast<const ulonglong2*>(a.p1 + (size_t)loc * row_words);
        ulonglong2* out = reinterpret_cast<ulonglong2*>(a.pres + (size_t)(a.res_idx[idx] + j) * row_words);
        const int Wv = Wp >> 1;
        const int vpos = (M == 2 && !flip) ? Wv : 0, vneg = (M == 2 && !flip) ? 0 : Wv;
#pragma unroll 4
        for (int k = lane; k < Wv; k += 32) {
          const ulonglong2 x = u[k], y = __ldg(v + vpos + k);
          __stcs(out + k, make_ulonglong2(x.x | y.x, x.y | y.y));
          if (M == 2) {
            const ulonglong2 xn = u[Wv + k], yn = __ldg(v + vneg + k);
            __stcs(out + Wv + k, make_ulonglong2(xn.x | yn.x, xn.y | yn.y));
          }
        }
      }
      // ---- true score -> top-K candidate (src/methods.h:90-94, 253-264) ----
      if (lane == 0) {
        double score;
        int cases, ctrls;
        unsigned tmax;
        if (M == 1) {
          cases = (int)(nc0[0] + ncn[0]);
          tmax = t0[0] + nd[0];
          ctrls = (int)tmax - cases;
          score = a.diagD[diag_base(tmax) + cases];
        } else {
          const unsigned tp = t0[0] + nd[0], tn = t0[M - 1] + nd[M - 1];
          const unsigned case_pos = nc0[0] + ncn[0], ctrl_neg = tp - case_pos;
          const unsigned ctrl_pos = nc0[M - 1] + ncn[M - 1], case_neg = tn - ctrl_pos;
          score = a.diagD[diag_base(tp) + case_pos] + a.diagD[diag_base(tn) + case_neg];
          cases = (int)(case_pos + case_neg);
          ctrls = (int)(ctrl_pos + ctrl_neg);
          tmax = max(tp, tn);
        }
        if (KEEP) atomicMax(a.max_total, tmax);
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
    __syncwarp();
  }
  // best[q]: register i = R*sub + (q >> 1), half q & 1  ->  permutation w*32 + ((i & 1)*2 + (q & 1))*8 + (i >> 1)
#pragma unroll
  for (int q = 0; q < 2 * R; q++) {
    const int i = R * sub + (q >> 1);
    const int r = w * 32 + ((i & 1) * 2 + (q & 1)) * 8 + (i >> 1);
    if (r < a.Ip && best[q] > 0.0f) atomicMax(a.perm_max + r, __float_as_int(best[q]));
  }
}

// few permutations, no count tables in play; GCRE_TEST_NO_SPLIT=1 (test hook) keeps the one-word-per-lane kernel
static inline bool sparse_sc_enabled(int Ip, int n_perm_blocks) {
  return Ip <= sparse_sc::MAX_PERMS && n_perm_blocks == 1 && std::getenv("GCRE_TEST_NO_SPLIT") == nullptr;
}
static inline bool sparse_sc_applies(const JoinParams& jp, const SparseParams& sp) {
  return sparse_sc_enabled(jp.Ip, sp.n_perm_blocks) && !sp.pcnt0 && !sp.pcnt1 && !sp.pcnt_res;
}

template <int M, bool KEEP, int NW>
static inline void launch_sparse_sc_ct(cudaStream_t stream, const JoinParams& jp, const SparseParams& sp, int sm_count) {
  const unsigned long long want = (sp.n_units + sparse_sc::WARPS - 1) / sparse_sc::WARPS;
  const unsigned grid = (unsigned)std::min<unsigned long long>(want, (unsigned long long)sm_count * sparse_sc::min_blocks(NW));
  if (sparse_wide(sp.n)) join_sparse_sc_kernel<M, KEEP, uint32_t, NW><<<grid, sparse_sc::THREADS, 0, stream>>>(jp, sp);
  else join_sparse_sc_kernel<M, KEEP, uint16_t, NW><<<grid, sparse_sc::THREADS, 0, stream>>>(jp, sp);
}

template <int M, bool KEEP>
static inline void launch_sparse_sc_nw(cudaStream_t stream, const JoinParams& jp, const SparseParams& sp, int sm_count) {
  if (jp.Ip <= 128) launch_sparse_sc_ct<M, KEEP, 4>(stream, jp, sp, sm_count);
  else if (jp.Ip <= 256) launch_sparse_sc_ct<M, KEEP, 8>(stream, jp, sp, sm_count);
  else launch_sparse_sc_ct<M, KEEP, 16>(stream, jp, sp, sm_count);
}

static inline cudaError_t launch_join_sparse_sc(cudaStream_t stream, const JoinParams& jp, const SparseParams& sp, int M, bool keep, int sm_count) {
  if (sp.n_units == 0) return cudaSuccess;
  if (M == 1) {
    if (keep) launch_sparse_sc_nw<1, true>(stream, jp, sp, sm_count);
    else launch_sparse_sc_nw<1, false>(stream, jp, sp, sm_count);
  } else {
    if (keep) launch_sparse_sc_nw<2, true>(stream, jp, sp, sm_count);
    else launch_sparse_sc_nw<2, false>(stream, jp, sp, sm_count);
  }
  return cudaGetLastError();
}

}  // namespace gcre
