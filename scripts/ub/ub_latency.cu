// Micro-benchmark: dependent-issue latency of DFMA on sm_100a and the independent-chain count
// needed to fill the fp64 pipe.  C chains per thread, W warps per SMSP.
#include <cstdio>
#include <cuda_runtime.h>
template <int C>
__global__ void k(double* out, int iters, double x) {
  double a[C];
  for (int i = 0; i < C; ++i) a[i] = threadIdx.x * 1e-3 + i;
  const double y = x * 0.5;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 16; ++rep)
#pragma unroll
      for (int i = 0; i < C; ++i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(x), "d"(y));
  }
  double s = 0;
  for (int i = 0; i < C; ++i) s += a[i];
  if (s == 1.2345) out[0] = s;
}
template <int C>
void run(double* d, int sms, int W) {
  const int iters = 1 << 12;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<C><<<sms, 128 * W>>>(d, 16, 1.0);
  cudaEventRecord(e0);
  k<C><<<sms, 128 * W>>>(d, iters, 1.0);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double cyc = ms * 1e-3 * clk * 1e3;
  const double per_round = cyc / (iters * 16.0);      // cycles per round of C DFMAs per warp
  printf("chains=%d warps/SMSP=%d: %.2f cycles per round (=%.2f per DFMA issued on the SMSP), %.1f%% of 2 cyc/inst\n", C, W,
         per_round, per_round / (C * W), 200.0 * C * W / per_round);
}
int main() {
  double* d; cudaMalloc(&d, 8);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int W = 1; W <= 4; ++W) { run<1>(d, sms, W); run<2>(d, sms, W); run<4>(d, sms, W); run<8>(d, sms, W); }
  return 0;
}
