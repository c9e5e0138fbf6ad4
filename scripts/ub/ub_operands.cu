// Micro-benchmark: does the register-operand count limit fp64 issue on sm_100a?
// 8 independent chains / thread, 4 warps / SMSP.  KIND 0: a = fma(a, x, y) (x, y shared -> reuse
// cache); 1: a = fma(a, b_i, y); 2: a = fma(b_i, c_i, a) (three distinct 64-bit registers);
// 3: a = fma(b_i, c_i, a) with b, c rotating so .reuse cannot apply; 4: a = a * b_i; 5: a = b_i * c_i + const-bank
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double cc[4] = {1.0000001, 0.5, 0.25, 3.0};
template <int KIND>
__global__ void k(double* out, int iters, double x) {
  double a[8], b[8], c[8];
  for (int i = 0; i < 8; ++i) { a[i] = 1.0 + threadIdx.x * 1e-3 + i; b[i] = 1.0 + 1e-9 * (i + threadIdx.x); c[i] = 1e-9 * i * x; }
  const double y = x * 0.5;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 8; ++rep)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (KIND == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(x), "d"(y));
        if (KIND == 1) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(b[i]), "d"(y));
        if (KIND == 2) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a[i]) : "d"(b[i]), "d"(c[i]));
        if (KIND == 3) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a[i]) : "d"(b[(i + rep) & 7]), "d"(c[(i + 3 * rep) & 7]));
        if (KIND == 4) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(a[i]) : "d"(b[i]));
        if (KIND == 5) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(b[i]), "d"(cc[rep & 3]));
      }
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += a[i] + b[i] + c[i];
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
  printf("%-34s %.2f SMSP cycles per warp instruction\n", nm, cyc / (iters * 64.0 * 4));
}
int main() {
  double* d; cudaMalloc(&d, 8);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  run<0>(d, sms, "fma(a, x, y)  shared x,y"); run<1>(d, sms, "fma(a, b_i, y)");
  run<2>(d, sms, "fma(b_i, c_i, a)"); run<3>(d, sms, "fma(b_j, c_k, a) rotating");
  run<4>(d, sms, "mul(a, b_i)"); run<5>(d, sms, "fma(a, b_i, const)");
  return 0;
}
