// Micro-benchmark: issue cost of the fp64 instruction kinds on sm_100a (8 independent chains per
// thread, 4 warps per SMSP): DFMA vs DMUL vs DADD vs DSETP+select vs mixed.
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND>
__global__ void k(double* out, int iters, double x) {
  double a[8];
  for (int i = 0; i < 8; ++i) a[i] = 1.0 + threadIdx.x * 1e-3 + i;
  const double y = x * 0.5;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 8; ++rep)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (KIND == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(x), "d"(y));
        if (KIND == 1) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(a[i]) : "d"(x));
        if (KIND == 2) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a[i]) : "d"(y));
        if (KIND == 3) asm volatile("{.reg .pred p; setp.gt.f64 p, %0, %1; selp.f64 %0, %0, %2, p;}" : "+d"(a[i]) : "d"(x), "d"(y));
        if (KIND == 4) {  // alternate DFMA / DMUL
          if (i & 1) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(a[i]) : "d"(x));
          else asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(x), "d"(y));
        }
        if (KIND == 5) {  // alternate DFMA / DADD
          if (i & 1) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a[i]) : "d"(y));
          else asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(x), "d"(y));
        }
      }
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 1.2345) out[0] = s;
}
template <int KIND>
void run(double* d, int sms, const char* nm) {
  const int iters = 1 << 12;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<KIND><<<sms, 512>>>(d, 16, 1.0);
  cudaEventRecord(e0);
  k<KIND><<<sms, 512>>>(d, iters, 1.0);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double cyc = ms * 1e-3 * clk * 1e3;
  printf("%-10s %.2f SMSP cycles per warp instruction\n", nm, cyc / (iters * 64.0 * 4));
}
int main() {
  double* d; cudaMalloc(&d, 8);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  run<0>(d, sms, "DFMA"); run<1>(d, sms, "DMUL"); run<2>(d, sms, "DADD"); run<3>(d, sms, "DSETP+SEL");
  run<4>(d, sms, "DFMA/DMUL"); run<5>(d, sms, "DFMA/DADD");
  return 0;
}
