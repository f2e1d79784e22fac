// ubench_banks.cu — which register pairs collide when one FFMA reads both?  (lead from round 1: pinning the NN
// scan's query registers to quad lanes changed its speed by 6 % with the same instruction stream.)
//
// Registers cannot be chosen from CUDA C, but a 128-bit shared-memory load lands in an ALIGNED register quad, so
// lane l of such a quad is a register with index = l (mod 4).  The kernel keeps 8 + 8 quads live and times
//     b = fma(a, c, b)        a = lane I of quad A_k,  b = lane J of quad B_k (updated in place),  c = a constant
// (c is served by the operand-reuse cache) for every (I, J): a 4 x 4 matrix of cycles per FFMA.  Equal rows / a
// checkerboard / a diagonal tell whether the banks are mod 2, mod 4, or something else.  Check the SASS
// (cuobjdump -sass) for the register numbers actually used before trusting a cell.
#include <cuda_runtime.h>
#include <stdio.h>
#define NQ 8
#define ITER 2048

template <int I, int J>
__global__ void __launch_bounds__(128) k(float* out, float c0) {
  __shared__ __align__(16) float4 stage[2 * NQ * 128];
  for (int i = threadIdx.x; i < 2 * NQ * 128; i += 128) stage[i] = make_float4(1.0f + i, 2.0f + i, 3.0f + i, 4.0f + i);
  __syncthreads();
  float a[NQ][4], b[NQ][4];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    unsigned int pa = (unsigned int)__cvta_generic_to_shared(&stage[q * 128 + threadIdx.x]);
    unsigned int pb = (unsigned int)__cvta_generic_to_shared(&stage[(NQ + q) * 128 + threadIdx.x]);
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a[q][0]), "=f"(a[q][1]), "=f"(a[q][2]), "=f"(a[q][3]) : "r"(pa));
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b[q][0]), "=f"(b[q][1]), "=f"(b[q][2]), "=f"(b[q][3]) : "r"(pb));
  }
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep)
#pragma unroll
      for (int q = 0; q < NQ; ++q) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(b[q][J]) : "f"(a[q][I]), "f"(c0));
  }
  float r = 0;
#pragma unroll
  for (int q = 0; q < NQ; ++q)
#pragma unroll
    for (int l = 0; l < 4; ++l) r += a[q][l] + b[q][l];
  out[blockIdx.x * 128 + threadIdx.x] = r;
}

template <int I, int J> float run(float* d) {
  const int blocks = 148 * 4;                         // 16 warps per SM, 4 per scheduler
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<I, J><<<blocks, 128>>>(d, 0.999f);
  cudaEventRecord(e0);
  k<I, J><<<blocks, 128>>>(d, 0.999f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ffma_per_smsp = (double)ITER * 4 * NQ * 4;      // 4 warps per scheduler
  return (float)(ms * 1e-3 * khz * 1e3 / ffma_per_smsp);
}

int main() {
  float* d; cudaMalloc(&d, 148 * 4 * 128 * 4);
  printf("cycles per FFMA (nominal clock), rows: lane of the multiplicand a, columns: lane of the accumulator b\n");
  float m[4][4];
#define ROW(I) m[I][0] = run<I, 0>(d); m[I][1] = run<I, 1>(d); m[I][2] = run<I, 2>(d); m[I][3] = run<I, 3>(d);
  ROW(0) ROW(1) ROW(2) ROW(3)
  for (int i = 0; i < 4; ++i) printf("a lane %d:  %.3f  %.3f  %.3f  %.3f\n", i, m[i][0], m[i][1], m[i][2], m[i][3]);
  return 0;
}
