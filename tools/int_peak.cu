// Integer / issue-rate microbenchmark for the rooflines of the join kernels ("replace the nominal POPC peak by a
// measured one"; the sparse kernels are LOP3 / issue bound, so their roof is the ALU pipe and the warp-instruction issue rate).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/int_peak tools/int_peak.cu && tools/int_peak > profiles/r2_int_peak.json
//
// Every kernel runs `CHAINS` independent dependency chains per thread of ONE instruction kind written as volatile inline PTX
// (the compiler can neither drop nor merge them; cuobjdump -sass of this file shows the expected opcode counts), on
// 148 x 8 CTAs of 256 threads, timed with CUDA events after a warm-up (best of 5).  The measured quantity is warp
// instructions per SECOND; the per-clock figures divide by the device's maximum SM clock (cudaDevAttrClockRate: 1,965 MHz on
// B200, which bench.py's NVML samples show the join kernels run at), so they are lower bounds if the clock dipped.
// "max.f32" chains are fused pairwise into FMNMX3 by ptxas: that line counts PTX operations, not SASS instructions.
// Output: one JSON object with warp-instructions per second and per clock per SM for each kind.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e = (x);                                                              \
    if (e != cudaSuccess) {                                                           \
      fprintf(stderr, "%s: %s (%s:%d)\n", #x, cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                        \
    }                                                                                 \
  } while (0)

constexpr int CHAINS = 8;
constexpr int UNROLL = 16;  // CHAINS * UNROLL measured instructions per loop trip; loop overhead: 2 instructions (IADD3 + BRA) per trip

enum Kind { POPC, LOP3, IADD3, IMAD, SHF, PRMT, FMNMX, LOP3_IMAD, LOP3_POPC, HS_MIX, N_KINDS };
static const char* kind_name[N_KINDS] = {"popc_b32", "lop3_b32", "iadd3", "imad", "shf", "prmt", "fmnmx", "lop3+imad 1:1", "lop3+popc 7:1",
                                         "lop3+imad+popc 6:1:1"};

template <int K>
__device__ __forceinline__ void op(uint32_t& r, uint32_t a, uint32_t b, int u) {
  if (K == POPC) asm volatile("popc.b32 %0, %0;" : "+r"(r));
  if (K == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r) : "r"(a), "r"(b));
  if (K == IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(r) : "r"(a));
  if (K == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r) : "r"(a), "r"(b));
  if (K == SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(r) : "r"(a), "r"(b));
  if (K == PRMT) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(r) : "r"(a), "r"(b));
  if (K == FMNMX) asm volatile("max.f32 %0, %0, %1;" : "+r"(r) : "r"(a));
  if (K == LOP3_IMAD) {
    if (u & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r) : "r"(a), "r"(b));
    else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r) : "r"(a), "r"(b));
  }
  if (K == LOP3_POPC) {
    if ((u & 7) == 7) asm volatile("popc.b32 %0, %0;" : "+r"(r));
    else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r) : "r"(a), "r"(b));
  }
  if (K == HS_MIX) {
    if ((u & 7) == 7) asm volatile("popc.b32 %0, %0;" : "+r"(r));
    else if ((u & 7) == 3) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r) : "r"(a), "r"(b));
    else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r) : "r"(a), "r"(b));
  }
}

template <int K>
__global__ void __launch_bounds__(256) bench_kernel(uint32_t* out, int trips, unsigned long long* clocks) {
  uint32_t r[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; c++) r[c] = threadIdx.x * 2654435761u + c * 40503u + blockIdx.x;
  const uint32_t a = threadIdx.x | 0x01010101u, b = blockIdx.x * 7 + 3;
  const unsigned long long t0 = clock64();
#pragma unroll 1
  for (int t = 0; t < trips; t++) {
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
#pragma unroll
      for (int c = 0; c < CHAINS; c++) op<K>(r[c], a, b, u * CHAINS + c);
    }
  }
  const unsigned long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) acc ^= r[c];
  if (acc == 0x12345678u) out[0] = acc;  // keeps the chains alive
  if (threadIdx.x == 0 && blockIdx.x == 0) clocks[0] = t1 - t0;
}

static double g_clock_hz = 1.965e9;

template <int K>
static void run(int sm_count, int ctas_per_sm, uint32_t* d_out, unsigned long long* d_clk, bool first) {
  const int trips = 4096, grid = sm_count * ctas_per_sm;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int w = 0; w < 3; w++) bench_kernel<K><<<grid, 256>>>(d_out, trips, d_clk);
  CK(cudaDeviceSynchronize());
  float best_ms = 1e30f;
  unsigned long long clk = 0;
  for (int rep = 0; rep < 5; rep++) {
    CK(cudaEventRecord(e0));
    bench_kernel<K><<<grid, 256>>>(d_out, trips, d_clk);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best_ms) {
      best_ms = ms;
      CK(cudaMemcpy(&clk, d_clk, 8, cudaMemcpyDeviceToHost));
    }
  }
  const double warp_inst = (double)grid * 8 /*warps per CTA*/ * trips * (double)(UNROLL * CHAINS);
  const double all_inst = warp_inst + (double)grid * 8 * trips * 2.0;
  const double sec = best_ms * 1e-3;
  (void)clk;
  const double clks = sec * g_clock_hz * sm_count;  // SM-clocks of the launch at the maximum SM clock
  printf("%s\n  {\"kind\": \"%s\", \"ctas_per_sm\": %d, \"ms\": %.4f, \"warp_inst_per_s\": %.4e, "
         "\"warp_inst_per_clk_per_sm\": %.3f, \"thread_ops_per_clk_per_sm\": %.1f, \"incl_loop_overhead_warp_inst_per_s\": %.4e}",
         first ? "" : ",", kind_name[K], ctas_per_sm, best_ms, warp_inst / sec, warp_inst / clks, 32.0 * warp_inst / clks, all_inst / sec);
}

int main() {
  int dev = 0, sm_count = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  sm_count = prop.multiProcessorCount;
  int khz = 0;
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
  g_clock_hz = khz * 1e3;
  uint32_t* d_out;
  unsigned long long* d_clk;
  CK(cudaMalloc(&d_out, 64));
  CK(cudaMalloc(&d_clk, 64));
  printf("{\"device\": \"%s\", \"sm_count\": %d, \"sm_clock_mhz_max\": %.0f, \"chains_per_thread\": %d, \"threads_per_cta\": 256, \"results\": [", prop.name,
         sm_count, g_clock_hz / 1e6, CHAINS);
  const int cps = 8;  // 8 CTAs x 8 warps = 64 warps per SM (16 per sub-partition)
  run<POPC>(sm_count, cps, d_out, d_clk, true);
  run<LOP3>(sm_count, cps, d_out, d_clk, false);
  run<IADD3>(sm_count, cps, d_out, d_clk, false);
  run<IMAD>(sm_count, cps, d_out, d_clk, false);
  run<SHF>(sm_count, cps, d_out, d_clk, false);
  run<PRMT>(sm_count, cps, d_out, d_clk, false);
  run<FMNMX>(sm_count, cps, d_out, d_clk, false);
  run<LOP3_IMAD>(sm_count, cps, d_out, d_clk, false);
  run<LOP3_POPC>(sm_count, cps, d_out, d_clk, false);
  run<HS_MIX>(sm_count, cps, d_out, d_clk, false);
  printf("\n]}\n");
  return 0;
}
