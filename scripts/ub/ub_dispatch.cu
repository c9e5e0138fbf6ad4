// Micro-benchmark: does a half-rate fp64 instruction occupy the SMSP dispatch port for 2 cycles?
// 8 independent DFMA chains per thread + NI independent ALU (LOP3) / FMA-pipe (IMAD) / FSEL ops
// per 8 DFMAs.  Reports cycles per (8 DFMA + NI other) group per SMSP-warp.
#include <cstdio>
#include <cuda_runtime.h>
template <int NI, int KIND>
__global__ void __launch_bounds__(256) k(double* out, int iters, double x, int seed) {
  double a[8];
  unsigned b[16];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int i = 0; i < 16; ++i) b[i] = threadIdx.x + i + seed;
  const double y = x * 0.5;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(x), "d"(y));
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        if (KIND == 0) asm volatile("xor.b32 %0, %0, %1;" : "+r"(b[i % 16]) : "r"(seed));
        else if (KIND == 1) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(b[i % 16]) : "r"(seed));
        else asm volatile("{.reg .pred p; setp.gt.u32 p, %0, %1; selp.b32 %0, %0, %1, p;}" : "+r"(b[i % 16]) : "r"(seed));
      }
    }
  }
  double s = 0; unsigned t = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  for (int i = 0; i < 16; ++i) t ^= b[i];
  if (s == 1.2345 || t == 0x12345) out[0] = s + t;
}
template <int NI, int KIND>
void run(double* d, int sms, const char* nm) {
  const int iters = 1 << 13, grid = sms * 2;  // 16 warps/SM = 4 per SMSP
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<NI, KIND><<<grid, 256>>>(d, 64, 1.0, 3);
  cudaEventRecord(e0);
  k<NI, KIND><<<grid, 256>>>(d, iters, 1.0, 3);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  // per SMSP: 4 warps x iters x 4 groups
  const double groups = 4.0 * iters * 4;
  const double cyc = ms * 1e-3 * clk * 1e3 / groups;
  printf("%-5s NI=%2d  %.3f ms  %.2f cycles per (8 DFMA + %d other) at nominal %d MHz  DFMA TFLOP/s %.2f\n", nm, NI, ms, cyc, NI, clk / 1000,
         2.0 * 8 * 4 * iters * 256.0 * grid / (ms * 1e-3) / 1e12);
}
int main() {
  double* d; cudaMalloc(&d, 8);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  run<0, 0>(d, sms, "none");
  run<4, 0>(d, sms, "lop3"); run<8, 0>(d, sms, "lop3"); run<16, 0>(d, sms, "lop3");
  run<4, 1>(d, sms, "imad"); run<8, 1>(d, sms, "imad"); run<16, 1>(d, sms, "imad");
  run<4, 2>(d, sms, "sel"); run<8, 2>(d, sms, "sel"); run<16, 2>(d, sms, "sel");
  return 0;
}
