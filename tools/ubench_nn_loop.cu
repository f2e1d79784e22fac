// ubench_nn_loop.cu — how fast can the NN filter's instruction mix issue when nothing else is in the way?
// Same per-chunk work as nn_scan_f32_kernel (8 queries per thread, 32 records per chunk, 4 records interleaved,
// 3 FFMA per evaluation) in four stages: FFMA only, + FMNMX3, + broadcast LDS.128 of the records, + the
// per-chunk bookkeeping.  Timed with clock64() inside the kernel (SM cycles, independent of the clock rate),
// 5 CTAs of 128 threads per SM like the library kernel.  Prints cycles per warp-evaluation per SM sub-partition;
// the library kernel measures 5.0 (ncu: 9.08 ms at C3).
#include <cuda_runtime.h>
#include <stdio.h>
#define Q 8
#define CHUNK 32
#define IL 4
#define NCHUNKS 2048
__device__ __forceinline__ float min3(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

template <int MODE>
__global__ void __launch_bounds__(128, 5) k(const float4* __restrict__ recs, float* out, long long* cyc, float seed) {
  __shared__ __align__(16) float4 tile[1024];
  for (int i = threadIdx.x; i < 1024; i += 128) tile[i] = recs[i];
  __syncthreads();
  float x[Q][3], best[Q], second[Q];
  int bchunk[Q];
#pragma unroll
  for (int k = 0; k < Q; ++k) { x[k][0] = seed + k + threadIdx.x * 1e-3f; x[k][1] = seed * 0.5f + k; x[k][2] = seed * 0.25f - k; best[k] = 3e38f; second[k] = 3e38f; bchunk[k] = 0; }
  float4 r0 = tile[0], r1 = tile[1], r2 = tile[2], r3 = tile[3];
  const long long t0 = clock64();
  for (int c = 0; c < NCHUNKS; ++c) {
    float cm[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) cm[k] = 3e38f;
    const float4* ch = &tile[(c & 31) * CHUNK];
#pragma unroll
    for (int j = 0; j < CHUNK; j += IL) {
      float4 a[IL];
      if (MODE >= 2) {
#pragma unroll
        for (int i = 0; i < IL; ++i) a[i] = ch[j + i];
      } else {
        a[0] = r0; a[1] = r1; a[2] = r2; a[3] = r3;
        r0.w += 1.0f; r1.w += 1.0f; r2.w += 1.0f; r3.w += 1.0f;        // keep the records loop-variant (4 FADD per 96 FFMA)
      }
      float v[IL][Q];
#pragma unroll
      for (int i = 0; i < IL; ++i)
#pragma unroll
        for (int k = 0; k < Q; ++k) v[i][k] = __fmaf_rn(x[k][2], a[i].z, a[i].w);
#pragma unroll
      for (int i = 0; i < IL; ++i)
#pragma unroll
        for (int k = 0; k < Q; ++k) v[i][k] = __fmaf_rn(x[k][1], a[i].y, v[i][k]);
#pragma unroll
      for (int i = 0; i < IL; ++i)
#pragma unroll
        for (int k = 0; k < Q; ++k) v[i][k] = __fmaf_rn(x[k][0], a[i].x, v[i][k]);
      if (MODE >= 1) {
#pragma unroll
        for (int i = 0; i < IL; i += 2)
#pragma unroll
          for (int k = 0; k < Q; ++k) cm[k] = min3(cm[k], v[i][k], v[i + 1][k]);
      } else {
#pragma unroll
        for (int k = 0; k < Q; ++k) cm[k] += v[0][k] * 0.0f + v[1][k] * 0.0f + v[2][k] * 0.0f + v[3][k] * 0.0f;   // placeholder use (not timed as a mode of interest)
      }
    }
    if (MODE >= 3) {
#pragma unroll
      for (int k = 0; k < Q; ++k) {
        second[k] = fminf(second[k], fmaxf(cm[k], best[k]));
        if (cm[k] < best[k]) { best[k] = cm[k]; bchunk[k] = c; }
      }
    } else {
#pragma unroll
      for (int k = 0; k < Q; ++k) best[k] = fminf(best[k], cm[k]);
    }
  }
  const long long t1 = clock64();
  float r = 0;
#pragma unroll
  for (int k = 0; k < Q; ++k) r += best[k] + second[k] + bchunk[k];
  out[blockIdx.x * 128 + threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char* name, const float4* recs, float* out, long long* cyc) {
  const int blocks = 148 * 5;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, 128>>>(recs, out, cyc, 1.0f);
  cudaEventRecord(e0);
  k<MODE><<<blocks, 128>>>(recs, out, cyc, 1.0f);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  long long* h = (long long*)malloc(blocks * sizeof(long long));
  cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
  double s = 0; for (int i = 0; i < blocks; ++i) s += (double)h[i];
  const double per_block = s / blocks;                       // cycles during which 5 CTAs x 4 warps shared the SM
  // one warp per CTA per sub-partition: 5 warps per SMSP, each doing NCHUNKS*CHUNK*Q warp-evaluations
  const double evals_per_smsp = 5.0 * NCHUNKS * CHUNK * Q;
  printf("%-44s %.3f cycles per warp-evaluation per SMSP by clock64(); %.3f by wall time (%.3f ms at %d MHz nominal)\n", name,
         per_block / evals_per_smsp, ms * 1e-3 * khz * 1e3 / evals_per_smsp, ms, khz / 1000);
  free(h);
}
int main() {
  float4* recs; float* out; long long* cyc;
  cudaMalloc(&recs, 1024 * sizeof(float4)); cudaMemset(recs, 0, 1024 * sizeof(float4));
  cudaMalloc(&out, 148 * 5 * 128 * 4); cudaMalloc(&cyc, 148 * 5 * 8);
  run<1>("3 FFMA + 1/2 FMNMX3 (records in registers)", recs, out, cyc);
  run<2>(" + broadcast LDS.128 per record", recs, out, cyc);
  run<3>(" + per-chunk bookkeeping (= the kernel's loop)", recs, out, cyc);
  return 0;
}
