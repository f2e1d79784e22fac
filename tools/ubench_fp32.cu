// ubench_fp32.cu — throughput of scalar vs packed FP32 instructions on sm_100a (one-off
// microbenchmark that decided the instruction mix of the NN scan).
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long u64;
#define ITER 4096
#define NCH 8
template <int MODE>
__global__ void k(float* out, float seed) {
  float a[NCH], b[NCH];
  u64 A[NCH], B[NCH];
  const float s = seed + threadIdx.x * 1e-9f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) { a[i] = s + i; b[i] = s * 0.5f + i; asm("mov.b64 %0, {%1, %2};" : "=l"(A[i]) : "f"(a[i]), "f"(b[i])); B[i] = A[i]; }
  const float c0 = s * 1.0001f, c1 = s * 0.9999f;
  u64 Cc; asm("mov.b64 %0, {%1, %2};" : "=l"(Cc) : "f"(c0), "f"(c1));
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      if (MODE == 0) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c0), "f"(c1)); }                       // FFMA
      if (MODE == 1) { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c0)); }                                     // FADD
      if (MODE == 2) { asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(A[i]) : "l"(Cc)); }                               // FFMA2 (2 distinct)
      if (MODE == 3) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(A[i]) : "l"(Cc)); }                                   // FADD2
      if (MODE == 4) { asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(A[i]) : "l"(Cc)); }                                   // FMUL2
      if (MODE == 5) { asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(A[i]) : "l"(B[i])); }                             // FFMA2 d = b*b + d
      if (MODE == 6) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(A[i]) : "l"(Cc)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c0), "f"(c1)); }   // FADD2 + FFMA
      if (MODE == 7) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(A[i]) : "l"(Cc)); asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c0), "f"(c1)); }      // FADD2 + FMNMX3
      if (MODE == 8) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c0), "f"(c1)); asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(b[i]) : "f"(c0), "f"(c1)); } // FFMA + FMNMX3
      if (MODE == 9) { asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(A[i]) : "l"(Cc)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c0), "f"(c1)); }  // FFMA2 + FFMA
      if (MODE == 10) { asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c0), "f"(c1)); }                           // FMNMX3
      if (MODE == 11) { asm volatile("min.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c0)); }                                        // FMNMX
      if (MODE == 12) { asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(A[i])); }                                          // FFMA2 (1 distinct)
      if (MODE == 13) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(A[i]) : "l"(Cc)); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c0)); }   // FADD2 + FADD
    }
  }
  float r = 0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(A[i])); r += a[i] + b[i] + lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* name, int per_iter, float* d) {
  const int blocks = 148 * 4, threads = 256;     // 32 warps / SM, 8 per scheduler
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(d, 1.0f);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(d, 1.0f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double warp_instr_per_smsp = (double)ITER * NCH * per_iter * (blocks * threads / 32) / (148.0 * 4);
  const double cycles = ms * 1e-3 * clk * 1e3;
  printf("%-28s %7.3f ms   %.3f cycles per warp-instruction per SMSP (at %d MHz nominal)\n", name, ms, cycles / warp_instr_per_smsp, clk / 1000);
}
int main() {
  float* d; cudaMalloc(&d, 148 * 4 * 256 * 4);
  run<0>("FFMA", 1, d); run<1>("FADD", 1, d); run<2>("FFMA2 a*c+c", 1, d); run<12>("FFMA2 a*a+a", 1, d); run<5>("FFMA2 b*b+a", 1, d);
  run<3>("FADD2", 1, d); run<4>("FMUL2", 1, d); run<10>("FMNMX3", 1, d); run<11>("FMNMX", 1, d);
  run<6>("FADD2+FFMA (per pair)", 2, d); run<13>("FADD2+FADD (per pair)", 2, d); run<7>("FADD2+FMNMX3 (per pair)", 2, d);
  run<8>("FFMA+FMNMX3 (per pair)", 2, d); run<9>("FFMA2+FFMA (per pair)", 2, d);
  return 0;
}
