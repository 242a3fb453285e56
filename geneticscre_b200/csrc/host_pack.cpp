// Host-side packing of R-format integer matrices (one int32 per patient, any non-zero = carrier, src/gcre_paths.h:65)
// into 64-bit words (patient c -> word c/64, bit c%64, src/gcre_paths.h:67) with a few host threads.
//
// Why on the host: the int matrices are 32x larger than the bits, and a level schedule uploads 1.2 GB of them per method
// over a 55 GB/s PCIe link (22 ms) - if the host cores can stream them faster than that, packing first and uploading the
// 37 MB of bits wins, and it also takes pageable R memory (which the driver stages at a fraction of the link rate) off
// the copy path.  The device-side pack kernel stays for small inputs and as the fallback (gcre_capi.cu picks).
#include <atomic>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace gcre_host {

static inline uint64_t pack64_scalar(const int32_t* p, int n) {
  uint64_t w = 0;
  for (int b = 0; b < n; b++) w |= (uint64_t)(p[b] != 0) << b;
  return w;
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) static void pack_rows_avx2(const int32_t* data, size_t r0, size_t r1, int cols, uint64_t* out, size_t out_stride) {
  const int full = cols / 64, tail = cols % 64;
  const __m256i zero = _mm256_setzero_si256();
  for (size_t r = r0; r < r1; r++) {
    const int32_t* src = data + r * (size_t)cols;
    uint64_t* dst = out + r * out_stride;
    for (int k = 0; k < full; k++) {
      uint64_t w = 0;
      for (int g = 0; g < 8; g++) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + k * 64 + g * 8));
        const unsigned m = (unsigned)_mm256_movemask_ps(_mm256_castsi256_ps(_mm256_cmpeq_epi32(v, zero)));
        w |= (uint64_t)((~m) & 0xffu) << (g * 8);
      }
      dst[k] = w;
    }
    if (tail) dst[full] = pack64_scalar(src + full * 64, tail);
  }
}
#endif

#if defined(__x86_64__)
// AVX-512: one compare gives the 16 bits of 16 patients (the AVX2 path needs compare + movemask per 8 and is bound by the cores,
// ~7 GB/s per thread; this one by the memory system)
__attribute__((target("avx512f"))) static void pack_rows_avx512(const int32_t* data, size_t r0, size_t r1, int cols, uint64_t* out, size_t out_stride) {
  const int full = cols / 64, tail = cols % 64;
  const __m512i zero = _mm512_setzero_si512();
  for (size_t r = r0; r < r1; r++) {
    const int32_t* src = data + r * (size_t)cols;
    uint64_t* dst = out + r * out_stride;
    for (int k = 0; k < full; k++) {
      const int32_t* p = src + k * 64;
      const uint64_t m0 = _mm512_cmpneq_epi32_mask(_mm512_loadu_si512(p), zero);
      const uint64_t m1 = _mm512_cmpneq_epi32_mask(_mm512_loadu_si512(p + 16), zero);
      const uint64_t m2 = _mm512_cmpneq_epi32_mask(_mm512_loadu_si512(p + 32), zero);
      const uint64_t m3 = _mm512_cmpneq_epi32_mask(_mm512_loadu_si512(p + 48), zero);
      dst[k] = m0 | (m1 << 16) | (m2 << 32) | (m3 << 48);
    }
    if (tail) dst[full] = pack64_scalar(src + full * 64, tail);
  }
}
#endif

static void pack_rows_scalar(const int32_t* data, size_t r0, size_t r1, int cols, uint64_t* out, size_t out_stride) {
  const int full = cols / 64, tail = cols % 64;
  for (size_t r = r0; r < r1; r++) {
    const int32_t* src = data + r * (size_t)cols;
    uint64_t* dst = out + r * out_stride;
    for (int k = 0; k < full; k++) dst[k] = pack64_scalar(src + k * 64, 64);
    if (tail) dst[full] = pack64_scalar(src + full * 64, tail);
  }
}

// out: rows x out_stride words, the first ceil(cols/64) of each row are written
void pack_i32_rows(const int32_t* data, size_t rows, int cols, uint64_t* out, size_t out_stride, int threads) {
  auto work = [&](size_t r0, size_t r1) {
#if defined(__x86_64__)
    static const bool no512 = [] { const char* e = std::getenv("GCRE_HOST_PACK_ISA"); return e && std::strcmp(e, "avx2") == 0; }();  // A/B knob
    if (!no512 && __builtin_cpu_supports("avx512f")) {
      pack_rows_avx512(data, r0, r1, cols, out, out_stride);
      return;
    }
    if (__builtin_cpu_supports("avx2")) {
      pack_rows_avx2(data, r0, r1, cols, out, out_stride);
      return;
    }
#endif
    pack_rows_scalar(data, r0, r1, cols, out, out_stride);
  };
  if (threads <= 1 || rows < 2 * (size_t)threads) {
    work(0, rows);
    return;
  }
  // Rows are handed out in small chunks from a shared counter (the calling thread takes part): on a shared host a thread whose
  // core is taken away for a few milliseconds then delays only its current chunk, not a sixteenth of the matrix - with a static
  // split single end-to-end steps were stretched by tens of milliseconds.
  const size_t chunk = std::max<size_t>(1, std::min<size_t>(64, rows / ((size_t)threads * 8)));
  std::atomic<size_t> next{0};
  auto worker = [&]() {
    for (;;) {
      const size_t r0 = next.fetch_add(chunk, std::memory_order_relaxed);
      if (r0 >= rows) return;
      work(r0, std::min(rows, r0 + chunk));
    }
  };
  std::vector<std::thread> pool;
  try {
    pool.reserve(threads - 1);
    for (int t = 1; t < threads; t++) pool.emplace_back(worker);
  } catch (...) {  // thread creation refused (resource limits): the threads that did start and this one finish the work
  }
  worker();
  for (auto& th : pool) th.join();
}

}  // namespace gcre_host
