// gumbel_kernels.cu — streaming Gumbel-softmax "soft nearest neighbour" (SURVEY.md §8 f-4;
// reference dICP/nn.py:43-70, "not yet integrated" upstream and off in the shipped config).
//
//   logits_ij = (-|x_i - y_j|^2 + G_ij) / tau,   G = -log(-log(U + eps) + eps),  U ~ uniform[0,1)
//   out_i     = sum_j softmax_j(logits_i) * y_j            (all S columns of y)
//
// The reference materialises the (N,n,m) probability tensor (1.07 GB per 16k pair).  Here one
// thread owns one query and streams the pair's targets through shared memory with an online
// softmax (running maximum, normaliser and S-vector accumulator), so memory is O(n + m).  The
// noise comes from a counter-based generator (Philox4x32-10 keyed by seed, counter = pair, query,
// target/4), which makes the forward reproducible and lets the two backward kernels regenerate
// exactly the same noise instead of storing it:
//   gumbel_bwd_query_kernel   per query:  dx_i  = sum_j dl_ij * (-2 (x_i - y_j) / tau)
//   gumbel_bwd_target_kernel  per target: dy_j  = sum_i p_ij g_i + dl_ij * 2 (x_i - y_j) / tau  (first 3 columns)
// with p_ij = exp(l_ij - lse_i), dl_ij = p_ij (g_i . y_j - g_i . out_i).
// FP32 CUDA-core + MUFU bound (2 logs, 1 exp, a quarter Philox per evaluation).
#include "common.cuh"

namespace dicp {

struct Philox4 { unsigned int x, y, z, w; };

__host__ __device__ inline Philox4 philox4x32_10(unsigned int c0, unsigned int c1, unsigned int c2, unsigned int c3, unsigned int k0,
                                                 unsigned int k1) {
  const unsigned int M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  for (int r = 0; r < 10; ++r) {
    const unsigned long long p0 = (unsigned long long)M0 * c0, p1 = (unsigned long long)M1 * c2;
    const unsigned int hi0 = (unsigned int)(p0 >> 32), lo0 = (unsigned int)p0, hi1 = (unsigned int)(p1 >> 32), lo1 = (unsigned int)p1;
    const unsigned int n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  Philox4 o;
  o.x = c0; o.y = c1; o.z = c2; o.w = c3;
  return o;
}

// Gumbel noise of the four targets 4*jq .. 4*jq+3 for query i of pair b (nn.py:61-62)
template <typename T>
__device__ __forceinline__ void gumbel4(unsigned long long seed, int b, int i, int jq, T eps, T* g) {
  const Philox4 r = philox4x32_10((unsigned int)jq, (unsigned int)i, (unsigned int)b, 0u, (unsigned int)seed, (unsigned int)(seed >> 32));
  const unsigned int v[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const T u = (T)(v[k] >> 8) * (T)(1.0 / 16777216.0);            // uniform [0,1), 24 bits like torch.rand (fp32)
    g[k] = -log(-log(u + eps) + eps);
  }
}

constexpr int GB_THREADS = 128;
constexpr int GB_TILE = 256;      // target rows per shared-memory tile (multiple of 4)

// forward: out [N,n,S], lse [N,n]
template <typename T, int S>
__global__ void __launch_bounds__(GB_THREADS)
gumbel_fwd_kernel(const T* __restrict__ x, const T* __restrict__ y, int n, int m, unsigned long long seed, T eps, T inv_tau,
                  T* __restrict__ out, T* __restrict__ lse) {
  __shared__ T tile[GB_TILE][S];
  const int b = blockIdx.y;
  const int i = blockIdx.x * GB_THREADS + threadIdx.x;
  const int ic = i < n ? i : n - 1;
  const T* xp = x + ((size_t)b * n + ic) * 3;
  const T x0 = xp[0], x1 = xp[1], x2 = xp[2];
  const T* yb = y + (size_t)b * m * S;
  T M = -INFINITY, L = T(0), acc[S];
#pragma unroll
  for (int c = 0; c < S; ++c) acc[c] = T(0);
  for (int start = 0; start < m; start += GB_TILE) {
    const int cnt = min(GB_TILE, m - start);
    __syncthreads();
    for (int e = threadIdx.x; e < cnt * S; e += GB_THREADS) (&tile[0][0])[e] = yb[(size_t)start * S + e];
    __syncthreads();
    for (int j4 = 0; j4 < cnt; j4 += 4) {
      T g[4];
      gumbel4<T>(seed, b, ic, (start + j4) >> 2, eps, g);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int j = j4 + k;
        if (j >= cnt) break;
        const T d0 = x0 - tile[j][0], d1 = x1 - tile[j][1], d2 = x2 - tile[j][2];
        const T l = (g[k] - (d0 * d0 + d1 * d1 + d2 * d2)) * inv_tau;          // nn.py:58-64
        if (l > M) {
          const T sc = exp(M - l);
          L = L * sc + T(1);
#pragma unroll
          for (int c = 0; c < S; ++c) acc[c] = acc[c] * sc + tile[j][c];
          M = l;
        } else {
          const T p = exp(l - M);
          L += p;
#pragma unroll
          for (int c = 0; c < S; ++c) acc[c] += p * tile[j][c];
        }
      }
    }
  }
  if (i < n) {
    const T invL = T(1) / L;
#pragma unroll
    for (int c = 0; c < S; ++c) out[((size_t)b * n + i) * S + c] = acc[c] * invL;    // nn.py:65-68
    lse[(size_t)b * n + i] = M + log(L);
  }
}

// backward w.r.t. the queries; also writes D_i = g_i . out_i for the target pass
template <typename T, int S>
__global__ void __launch_bounds__(GB_THREADS)
gumbel_bwd_query_kernel(const T* __restrict__ x, const T* __restrict__ y, const T* __restrict__ out, const T* __restrict__ lse,
                        const T* __restrict__ gout, int n, int m, unsigned long long seed, T eps, T inv_tau,
                        T* __restrict__ gx, T* __restrict__ gdot) {
  __shared__ T tile[GB_TILE][S];
  const int b = blockIdx.y;
  const int i = blockIdx.x * GB_THREADS + threadIdx.x;
  const int ic = i < n ? i : n - 1;
  const size_t o = (size_t)b * n + ic;
  const T x0 = x[o * 3], x1 = x[o * 3 + 1], x2 = x[o * 3 + 2];
  T g[S], D = T(0);
#pragma unroll
  for (int c = 0; c < S; ++c) { g[c] = gout[o * S + c]; D += g[c] * out[o * S + c]; }
  const T ls = lse[o];
  const T* yb = y + (size_t)b * m * S;
  T a0 = T(0), a1 = T(0), a2 = T(0);
  for (int start = 0; start < m; start += GB_TILE) {
    const int cnt = min(GB_TILE, m - start);
    __syncthreads();
    for (int e = threadIdx.x; e < cnt * S; e += GB_THREADS) (&tile[0][0])[e] = yb[(size_t)start * S + e];
    __syncthreads();
    for (int j4 = 0; j4 < cnt; j4 += 4) {
      T gm[4];
      gumbel4<T>(seed, b, ic, (start + j4) >> 2, eps, gm);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int j = j4 + k;
        if (j >= cnt) break;
        const T d0 = x0 - tile[j][0], d1 = x1 - tile[j][1], d2 = x2 - tile[j][2];
        const T p = exp((gm[k] - (d0 * d0 + d1 * d1 + d2 * d2)) * inv_tau - ls);
        T gy = T(0);
#pragma unroll
        for (int c = 0; c < S; ++c) gy += g[c] * tile[j][c];
        const T dl = p * (gy - D);
        a0 += dl * d0; a1 += dl * d1; a2 += dl * d2;
      }
    }
  }
  if (i < n) {
    const T f = T(-2) * inv_tau;
    gx[o * 3] = f * a0; gx[o * 3 + 1] = f * a1; gx[o * 3 + 2] = f * a2;
    gdot[o] = D;
  }
}

// backward w.r.t. the targets: one thread per target, the queries stream through shared memory
template <typename T, int S>
__global__ void __launch_bounds__(GB_THREADS)
gumbel_bwd_target_kernel(const T* __restrict__ x, const T* __restrict__ y, const T* __restrict__ lse, const T* __restrict__ gout,
                         const T* __restrict__ gdot, int n, int m, unsigned long long seed, T eps, T inv_tau, T* __restrict__ gy_out) {
  constexpr int W = 3 + S + 2;     // x (3), g (S), lse, D
  __shared__ T tile[GB_TILE][W];
  const int b = blockIdx.y;
  const int j = blockIdx.x * GB_THREADS + threadIdx.x;
  const int jc = j < m ? j : m - 1;
  const T* yp = y + ((size_t)b * m + jc) * S;
  T yv[S], acc[S];
#pragma unroll
  for (int c = 0; c < S; ++c) { yv[c] = yp[c]; acc[c] = T(0); }
  for (int start = 0; start < n; start += GB_TILE) {
    const int cnt = min(GB_TILE, n - start);
    __syncthreads();
    for (int e = threadIdx.x; e < cnt; e += GB_THREADS) {
      const size_t o = (size_t)b * n + start + e;
      tile[e][0] = x[o * 3]; tile[e][1] = x[o * 3 + 1]; tile[e][2] = x[o * 3 + 2];
#pragma unroll
      for (int c = 0; c < S; ++c) tile[e][3 + c] = gout[o * S + c];
      tile[e][3 + S] = lse[o];
      tile[e][4 + S] = gdot[o];
    }
    __syncthreads();
    for (int e = 0; e < cnt; ++e) {
      T gm[4];
      gumbel4<T>(seed, b, start + e, jc >> 2, eps, gm);
      const T gk = gm[jc & 3];
      const T d0 = tile[e][0] - yv[0], d1 = tile[e][1] - yv[1], d2 = tile[e][2] - yv[2];
      const T p = exp((gk - (d0 * d0 + d1 * d1 + d2 * d2)) * inv_tau - tile[e][3 + S]);
      T gy = T(0);
#pragma unroll
      for (int c = 0; c < S; ++c) gy += tile[e][3 + c] * yv[c];
      const T dl = p * (gy - tile[e][4 + S]) * (T(2) * inv_tau);
#pragma unroll
      for (int c = 0; c < S; ++c) acc[c] += p * tile[e][3 + c];
      acc[0] += dl * d0; acc[1] += dl * d1; acc[2] += dl * d2;
    }
  }
  if (j < m) {
#pragma unroll
    for (int c = 0; c < S; ++c) gy_out[((size_t)b * m + j) * S + c] = acc[c];
  }
}

// debug / test helper: the noise tensor G [N,n,m] the kernels use
template <typename T>
__global__ void gumbel_noise_kernel(int n, int m, unsigned long long seed, T eps, T* __restrict__ G) {
  const int b = blockIdx.z, i = blockIdx.y;
  const int jq = blockIdx.x * blockDim.x + threadIdx.x;
  if (4 * jq >= m) return;
  T g[4];
  gumbel4<T>(seed, b, i, jq, eps, g);
  for (int k = 0; k < 4 && 4 * jq + k < m; ++k) G[((size_t)b * n + i) * m + 4 * jq + k] = g[k];
}

template <typename T>
static int gumbel_fwd_impl(const T* x, const T* y, int N, int n, int m, int S, unsigned long long seed, double eps, double tau, T* out,
                           T* lse, void* stream) {
  DICP_CHECK_ARG(x && y && out && lse, "null pointer");
  DICP_CHECK_ARG(N >= 1 && N <= 65535 && n >= 1 && m >= 1 && (S == 3 || S == 6) && tau > 0.0, "bad arguments");
  dim3 grid((n + GB_THREADS - 1) / GB_THREADS, N);
  cudaStream_t st = (cudaStream_t)stream;
  if (S == 3) gumbel_fwd_kernel<T, 3><<<grid, GB_THREADS, 0, st>>>(x, y, n, m, seed, (T)eps, (T)(1.0 / tau), out, lse);
  else gumbel_fwd_kernel<T, 6><<<grid, GB_THREADS, 0, st>>>(x, y, n, m, seed, (T)eps, (T)(1.0 / tau), out, lse);
  DICP_LAUNCH_CHECK();
  return 0;
}

template <typename T>
static int gumbel_bwd_impl(const T* x, const T* y, const T* out, const T* lse, const T* gout, int N, int n, int m, int S,
                           unsigned long long seed, double eps, double tau, T* gx, T* gy, T* gdot_ws, void* stream) {
  DICP_CHECK_ARG(x && y && out && lse && gout && gx && gy && gdot_ws, "null pointer");
  DICP_CHECK_ARG(N >= 1 && N <= 65535 && n >= 1 && m >= 1 && (S == 3 || S == 6) && tau > 0.0, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 gq((n + GB_THREADS - 1) / GB_THREADS, N), gt((m + GB_THREADS - 1) / GB_THREADS, N);
  const T e = (T)eps, it = (T)(1.0 / tau);
  if (S == 3) {
    gumbel_bwd_query_kernel<T, 3><<<gq, GB_THREADS, 0, st>>>(x, y, out, lse, gout, n, m, seed, e, it, gx, gdot_ws);
    gumbel_bwd_target_kernel<T, 3><<<gt, GB_THREADS, 0, st>>>(x, y, lse, gout, gdot_ws, n, m, seed, e, it, gy);
  } else {
    gumbel_bwd_query_kernel<T, 6><<<gq, GB_THREADS, 0, st>>>(x, y, out, lse, gout, n, m, seed, e, it, gx, gdot_ws);
    gumbel_bwd_target_kernel<T, 6><<<gt, GB_THREADS, 0, st>>>(x, y, lse, gout, gdot_ws, n, m, seed, e, it, gy);
  }
  DICP_LAUNCH_CHECK();
  return 0;
}

}  // namespace dicp

extern "C" {
int dicp_gumbel_nn_f32(const float* x, const float* y, int N, int n, int m, int S, unsigned long long seed, double eps, double tau,
                       float* out, float* lse, void* stream) {
  return dicp::gumbel_fwd_impl<float>(x, y, N, n, m, S, seed, eps, tau, out, lse, stream);
}
int dicp_gumbel_nn_f64(const double* x, const double* y, int N, int n, int m, int S, unsigned long long seed, double eps, double tau,
                       double* out, double* lse, void* stream) {
  return dicp::gumbel_fwd_impl<double>(x, y, N, n, m, S, seed, eps, tau, out, lse, stream);
}
int dicp_gumbel_nn_bwd_f32(const float* x, const float* y, const float* out, const float* lse, const float* grad_out, int N, int n, int m,
                           int S, unsigned long long seed, double eps, double tau, float* grad_x, float* grad_y, float* ws_n,
                           void* stream) {
  return dicp::gumbel_bwd_impl<float>(x, y, out, lse, grad_out, N, n, m, S, seed, eps, tau, grad_x, grad_y, ws_n, stream);
}
int dicp_gumbel_nn_bwd_f64(const double* x, const double* y, const double* out, const double* lse, const double* grad_out, int N, int n,
                           int m, int S, unsigned long long seed, double eps, double tau, double* grad_x, double* grad_y, double* ws_n,
                           void* stream) {
  return dicp::gumbel_bwd_impl<double>(x, y, out, lse, grad_out, N, n, m, S, seed, eps, tau, grad_x, grad_y, ws_n, stream);
}
int dicp_gumbel_noise_f32(int N, int n, int m, unsigned long long seed, double eps, float* G, void* stream) {
  DICP_CHECK_ARG(G && N >= 1 && N <= 65535 && n >= 1 && n <= 65535 && m >= 1, "bad arguments");
  dim3 grid(((m + 3) / 4 + 127) / 128, n, N);
  dicp::gumbel_noise_kernel<float><<<grid, 128, 0, (cudaStream_t)stream>>>(n, m, seed, (float)eps, G);
  DICP_LAUNCH_CHECK();
  return 0;
}
int dicp_gumbel_noise_f64(int N, int n, int m, unsigned long long seed, double eps, double* G, void* stream) {
  DICP_CHECK_ARG(G && N >= 1 && N <= 65535 && n >= 1 && n <= 65535 && m >= 1, "bad arguments");
  dim3 grid(((m + 3) / 4 + 127) / 128, n, N);
  dicp::gumbel_noise_kernel<double><<<grid, 128, 0, (cudaStream_t)stream>>>(n, m, seed, eps, G);
  DICP_LAUNCH_CHECK();
  return 0;
}
}
