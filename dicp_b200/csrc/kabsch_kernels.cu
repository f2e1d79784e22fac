// kabsch_kernels.cu — the per-point half of the Kabsch / SVD point-to-point step
// (ICP.py:557-563): centroids of the source points and of their nearest neighbours and the
// 3x3 cross-covariance, as ONE fused pass over the correspondences:
//   sums[b] = ( sum p , sum q , sum q p^T ),   q_i = target[b, idx[b,i], :3]
// from which  p_mean = Sp/n, q_mean = Sq/n, W = M/n - q_mean p_mean^T  (host side, 3x3).
// HBM bound: 12 (p) + 4 (idx) + 12 (gathered q) = 28 B per correspondence.  The 3x3 SVD, the
// determinant fix and the composition (ICP.py:566-578) are O(1) per pair and stay in torch.
// The adjoint kernel gives autograd the gradients to the points and (scatter-add) to the target.
#include "common.cuh"

namespace dicp {

constexpr int KB_THREADS = 128;
constexpr int KB_NSUM = 16;     // 3 + 3 + 9 (+1 pad)

template <typename T>
__global__ void __launch_bounds__(KB_THREADS)
kabsch_accum_kernel(const T* __restrict__ source, const T* __restrict__ target, const int32_t* __restrict__ idx, int n, int m, int S,
                    double* __restrict__ partials) {
  __shared__ double smem[(KB_THREADS / 32) * KB_NSUM];
  constexpr int U = 4;
  const int b = blockIdx.y;
  const T* src_b = source + (size_t)b * n * 3;
  const T* tgt_b = target + (size_t)b * m * S;
  const int32_t* idx_b = idx + (size_t)b * n;
  T acc[KB_NSUM];
#pragma unroll
  for (int i = 0; i < KB_NSUM; ++i) acc[i] = T(0);
  for (int base = blockIdx.x * (KB_THREADS * U); base < n; base += gridDim.x * (KB_THREADS * U)) {
    int jj[U];
    bool ok[U];
    T p[U][3], q[U][3];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = base + u * KB_THREADS + threadIdx.x;
      ok[u] = i < n;
      const int ic = ok[u] ? i : n - 1;
      jj[u] = idx_b[ic];
      p[u][0] = src_b[3 * ic]; p[u][1] = src_b[3 * ic + 1]; p[u][2] = src_b[3 * ic + 2];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const T* row = tgt_b + (size_t)jj[u] * S;
      q[u][0] = row[0]; q[u][1] = row[1]; q[u][2] = row[2];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!ok[u]) continue;
#pragma unroll
      for (int c = 0; c < 3; ++c) { acc[c] += p[u][c]; acc[3 + c] += q[u][c]; }
#pragma unroll
      for (int rr = 0; rr < 3; ++rr)
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[6 + rr * 3 + c] += q[u][rr] * p[u][c];
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int v = 0; v < KB_NSUM; ++v) {
    T x = acc[v];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (lane == 0) smem[warp * KB_NSUM + v] = (double)x;
  }
  __syncthreads();
  if (threadIdx.x < KB_NSUM) {
    double s = 0.0;
    for (int w = 0; w < KB_THREADS / 32; ++w) s += smem[w * KB_NSUM + threadIdx.x];
    partials[((size_t)b * gridDim.x + blockIdx.x) * KB_NSUM + threadIdx.x] = s;
  }
}

template <typename T>
__global__ void kabsch_finish_kernel(const double* __restrict__ partials, int nblk, T* __restrict__ sums) {
  const int b = blockIdx.x, v = threadIdx.x;
  if (v >= KB_NSUM) return;
  double s = 0.0;
  for (int j = 0; j < nblk; ++j) s += partials[((size_t)b * nblk + j) * KB_NSUM + v];     // fixed order: deterministic
  sums[(size_t)b * KB_NSUM + v] = (T)s;
}

// adjoint: pbar_i = gSp + gM^T q_i ; qbar_i = gSq + gM p_i -> atomicAdd into grad_target[idx_i, :3]
template <typename T>
__global__ void __launch_bounds__(KB_THREADS)
kabsch_bwd_kernel(const T* __restrict__ source, const T* __restrict__ target, const int32_t* __restrict__ idx, int n, int m, int S,
                  const T* __restrict__ grad_sums, T* __restrict__ grad_source, T* __restrict__ grad_target) {
  const int b = blockIdx.y;
  T g[KB_NSUM];
#pragma unroll
  for (int i = 0; i < KB_NSUM; ++i) g[i] = grad_sums[(size_t)b * KB_NSUM + i];
  const T* src_b = source + (size_t)b * n * 3;
  const T* tgt_b = target + (size_t)b * m * S;
  T* gs_b = grad_source + (size_t)b * n * 3;
  T* gt_b = grad_target + (size_t)b * m * S;
  for (int i = blockIdx.x * KB_THREADS + threadIdx.x; i < n; i += gridDim.x * KB_THREADS) {
    const int j = idx[(size_t)b * n + i];
    const T p[3] = {src_b[3 * i], src_b[3 * i + 1], src_b[3 * i + 2]};
    const T* row = tgt_b + (size_t)j * S;
    const T q[3] = {row[0], row[1], row[2]};
#pragma unroll
    for (int c = 0; c < 3; ++c) gs_b[3 * i + c] = g[c] + g[6 + c] * q[0] + g[9 + c] * q[1] + g[12 + c] * q[2];
#pragma unroll
    for (int rr = 0; rr < 3; ++rr)
      atomicAdd(gt_b + (size_t)j * S + rr, g[3 + rr] + g[6 + rr * 3] * p[0] + g[7 + rr * 3] * p[1] + g[8 + rr * 3] * p[2]);
  }
}

static int kb_blocks(int n) {
  int nb = (n + KB_THREADS * 8 - 1) / (KB_THREADS * 8);
  return nb < 1 ? 1 : (nb > 128 ? 128 : nb);
}

template <typename T>
static int kabsch_sums_impl(const T* source, const T* target, const int32_t* idx, int N, int n, int m, int S, T* sums, void* ws,
                            size_t ws_bytes, void* stream) {
  DICP_CHECK_ARG(source && target && idx && sums, "null pointer");
  DICP_CHECK_ARG(N >= 1 && N <= 65535 && n >= 1 && m >= 1 && (S == 3 || S == 6), "bad sizes");
  const int nblk = kb_blocks(n);
  Arena ar(ws, ws_bytes);
  double* partials = ar.take<double>((size_t)N * nblk * KB_NSUM);
  if (ws == nullptr || partials == nullptr) { set_error("dicp_kabsch_sums: workspace too small"); return DICP_ERR_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  kabsch_accum_kernel<T><<<dim3(nblk, N), KB_THREADS, 0, st>>>(source, target, idx, n, m, S, partials);
  DICP_LAUNCH_CHECK();
  kabsch_finish_kernel<T><<<N, 32, 0, st>>>(partials, nblk, sums);
  DICP_LAUNCH_CHECK();
  return 0;
}

template <typename T>
static int kabsch_bwd_impl(const T* source, const T* target, const int32_t* idx, int N, int n, int m, int S, const T* grad_sums,
                           T* grad_source, T* grad_target, void* stream) {
  DICP_CHECK_ARG(source && target && idx && grad_sums && grad_source && grad_target, "null pointer");
  DICP_CHECK_ARG(N >= 1 && N <= 65535 && n >= 1 && m >= 1 && (S == 3 || S == 6), "bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  DICP_CUDA(cudaMemsetAsync(grad_target, 0, (size_t)N * m * S * sizeof(T), st));
  kabsch_bwd_kernel<T><<<dim3(kb_blocks(n) * 2, N), KB_THREADS, 0, st>>>(source, target, idx, n, m, S, grad_sums, grad_source, grad_target);
  DICP_LAUNCH_CHECK();
  return 0;
}

}  // namespace dicp

extern "C" {
size_t dicp_kabsch_workspace_bytes(int N, int n) { return dicp::align_up((size_t)N * dicp::kb_blocks(n) * dicp::KB_NSUM * sizeof(double)); }
int dicp_kabsch_sums_f32(const float* source, const float* target, const int32_t* idx, int N, int n, int m, int S, float* sums, void* ws,
                         size_t ws_bytes, void* stream) {
  return dicp::kabsch_sums_impl<float>(source, target, idx, N, n, m, S, sums, ws, ws_bytes, stream);
}
int dicp_kabsch_sums_f64(const double* source, const double* target, const int32_t* idx, int N, int n, int m, int S, double* sums, void* ws,
                         size_t ws_bytes, void* stream) {
  return dicp::kabsch_sums_impl<double>(source, target, idx, N, n, m, S, sums, ws, ws_bytes, stream);
}
int dicp_kabsch_sums_bwd_f32(const float* source, const float* target, const int32_t* idx, int N, int n, int m, int S,
                             const float* grad_sums, float* grad_source, float* grad_target, void* stream) {
  return dicp::kabsch_bwd_impl<float>(source, target, idx, N, n, m, S, grad_sums, grad_source, grad_target, stream);
}
int dicp_kabsch_sums_bwd_f64(const double* source, const double* target, const int32_t* idx, int N, int n, int m, int S,
                             const double* grad_sums, double* grad_source, double* grad_target, void* stream) {
  return dicp::kabsch_bwd_impl<double>(source, target, idx, N, n, m, S, grad_sums, grad_source, grad_target, stream);
}
}
