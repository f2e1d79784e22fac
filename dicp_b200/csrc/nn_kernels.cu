// nn_kernels.cu — batched brute-force nearest neighbour (replaces cdist+argmin, nn.py:32-35).
//
// Layout: the targets of every pair are repacked once per call into 16-byte (x,y,z,0)
// records, padded with far-away sentinels to a multiple of NN_CHUNK.  A CTA owns
// THREADS*Q queries of one pair (Q queries live in the registers of each thread) and
// streams the pair's target records through shared memory in double-buffered cp.async
// tiles; every thread reads each record with one broadcast LDS.128 and evaluates Q
// direct-form squared distances against it (FP32 CUDA-core work, no tensor cores: a K=3
// distance is not a contraction).  Per chunk of NN_CHUNK records only a running minimum is
// kept (one FMNMX per evaluation); the chunk id is recorded when the chunk minimum beats
// the best so far, and the exact index is recovered afterwards by re-evaluating that one
// chunk with the bit-identical instruction sequence.  First-minimum tie-break like
// torch.argmin.  When N*n is too small to fill 148 SMs the target range is split over
// blockIdx.z and the partial (d2, idx) results are merged in split order.
#include "common.cuh"

namespace dicp {

template <typename T> struct Sentinel;
template <> struct Sentinel<float> { static __host__ __device__ float far() { return 1.0e18f; } static __device__ float inf() { return __int_as_float(0x7f800000); } };
template <> struct Sentinel<double> { static __host__ __device__ double far() { return 1.0e150; } static __device__ double inf() { return __longlong_as_double(0x7ff0000000000000LL); } };

template <typename T> __device__ __forceinline__ T t_min(T a, T b);
template <> __device__ __forceinline__ float t_min<float>(float a, float b) { return fminf(a, b); }
template <> __device__ __forceinline__ double t_min<double>(double a, double b) { return fmin(a, b); }

template <typename T>
__global__ void nn_pack_kernel(const T* __restrict__ tgt, int m, int m_pad, int S, typename Vec4<T>::type* __restrict__ packed) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m_pad) return;
  typename Vec4<T>::type v;
  if (j < m) {
    const T* row = tgt + ((size_t)b * m + j) * S;
    v.x = row[0]; v.y = row[1]; v.z = row[2]; v.w = T(0);
  } else {
    v.x = v.y = v.z = Sentinel<T>::far(); v.w = T(0);
  }
  packed[(size_t)b * m_pad + j] = v;
}

template <typename T>
int nn_pack_targets(const T* tgt, int N, int m, int S, typename Vec4<T>::type* packed, cudaStream_t st) {
  const int m_pad = nn_m_pad(m);
  dim3 grid((m_pad + 255) / 256, N);
  nn_pack_kernel<T><<<grid, 256, 0, st>>>(tgt, m, m_pad, S, packed);
  DICP_LAUNCH_CHECK();
  return 0;
}
template int nn_pack_targets<float>(const float*, int, int, int, float4*, cudaStream_t);
template int nn_pack_targets<double>(const double*, int, int, int, double4*, cudaStream_t);

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

constexpr int NN_THREADS = 128;
constexpr int NN_TILE = 512;      // records per shared-memory stage

template <typename T, int Q>
__global__ void __launch_bounds__(NN_THREADS)
nn_scan_kernel(const T* __restrict__ q, const double* __restrict__ state,
               const typename Vec4<T>::type* __restrict__ packed, int n, int m_pad, int tiles_per_split,
               int32_t* __restrict__ idx_out, T* __restrict__ d2_out, T* __restrict__ part_d2,
               int32_t* __restrict__ part_idx, const int* __restrict__ stop_flag) {
  typedef typename Vec4<T>::type V4;
  if (stop_flag != nullptr && *stop_flag != 0) return;
  __shared__ __align__(16) V4 tile[2][NN_TILE];

  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int split = blockIdx.z;
  const int N = gridDim.y;
  const V4* tg = packed + (size_t)b * m_pad;

  // this thread's queries, transformed into the target frame (ICP.py:137)
  T C[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, r[3] = {0, 0, 0};
  if (state != nullptr) {
    const double* rec = state + (size_t)b * ST_STRIDE;
#pragma unroll
    for (int i = 0; i < 9; ++i) C[i] = (T)rec[ST_C + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) r[i] = (T)rec[ST_R + i];
  }
  T x[Q][3], best[Q];
  int bchunk[Q];
  const int q0 = blockIdx.x * (NN_THREADS * Q) + tid;
#pragma unroll
  for (int k = 0; k < Q; ++k) {
    int i = q0 + k * NN_THREADS;
    if (i >= n) i = n - 1;
    const T* p = q + ((size_t)b * n + i) * 3;
    T pv[3] = {p[0], p[1], p[2]};
    if (state != nullptr) {
      T a[3];
      rotate(C, pv, a);
      x[k][0] = a[0] + r[0]; x[k][1] = a[1] + r[1]; x[k][2] = a[2] + r[2];
    } else {
      x[k][0] = pv[0]; x[k][1] = pv[1]; x[k][2] = pv[2];
    }
    best[k] = Sentinel<T>::inf();
    bchunk[k] = 0;
  }

  const int tiles_total = (m_pad + NN_TILE - 1) / NN_TILE;
  const int t_begin = split * tiles_per_split;
  const int t_end = min(t_begin + tiles_per_split, tiles_total);
  constexpr int UNITS = sizeof(V4) / 16;   // 16-byte cp.async units per record

  auto issue_tile = [&](int t, int buf) {
    const int start = t * NN_TILE;
    const int cnt = min(NN_TILE, m_pad - start);
    const char* src = (const char*)(tg + start);
    char* dst = (char*)&tile[buf][0];
    for (int u = tid; u < cnt * UNITS; u += NN_THREADS) cp_async16(dst + (size_t)u * 16, src + (size_t)u * 16);
    cp_async_commit();
  };

  if (t_begin < t_end) issue_tile(t_begin, 0);
  for (int t = t_begin; t < t_end; ++t) {
    const int buf = (t - t_begin) & 1;
    if (t + 1 < t_end) { issue_tile(t + 1, buf ^ 1); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncthreads();
    const int start = t * NN_TILE;
    const int nchunks = min(NN_TILE, m_pad - start) / NN_CHUNK;
    for (int c = 0; c < nchunks; ++c) {
      T cm[Q];
#pragma unroll
      for (int k = 0; k < Q; ++k) cm[k] = Sentinel<T>::inf();
#pragma unroll
      for (int j = 0; j < NN_CHUNK; ++j) {
        const V4 y = tile[buf][c * NN_CHUNK + j];
#pragma unroll
        for (int k = 0; k < Q; ++k) cm[k] = t_min<T>(cm[k], dist2<T>(x[k][0], x[k][1], x[k][2], y.x, y.y, y.z));
      }
#pragma unroll
      for (int k = 0; k < Q; ++k) {
        if (cm[k] < best[k]) { best[k] = cm[k]; bchunk[k] = start / NN_CHUNK + c; }
      }
    }
    __syncthreads();
  }

  // refine: the exact index inside the winning chunk (bit-identical arithmetic)
#pragma unroll
  for (int k = 0; k < Q; ++k) {
    const int i = q0 + k * NN_THREADS;
    if (i >= n) continue;
    int found = -1;
    const V4* ch = tg + (size_t)bchunk[k] * NN_CHUNK;
    if (t_begin < t_end) {
#pragma unroll 4
      for (int j = 0; j < NN_CHUNK; ++j) {
        const V4 y = ch[j];
        const T d = dist2<T>(x[k][0], x[k][1], x[k][2], y.x, y.y, y.z);
        if (found < 0 && d == best[k]) found = bchunk[k] * NN_CHUNK + j;
      }
    }
    if (found < 0) found = 0;
    const size_t o = (size_t)b * n + i;
    if (part_d2 == nullptr) {
      idx_out[o] = found;
      if (d2_out != nullptr) d2_out[o] = best[k];
    } else {
      part_d2[((size_t)split * N) * n + o] = best[k];
      part_idx[((size_t)split * N) * n + o] = found;
    }
  }
}

template <typename T>
__global__ void nn_merge_kernel(const T* __restrict__ part_d2, const int32_t* __restrict__ part_idx, size_t total, int nsplit,
                                int32_t* __restrict__ idx_out, T* __restrict__ d2_out, const int* __restrict__ stop_flag) {
  if (stop_flag != nullptr && *stop_flag != 0) return;
  const size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= total) return;
  T bd = part_d2[o];
  int bi = part_idx[o];
  for (int s = 1; s < nsplit; ++s) {
    const T d = part_d2[(size_t)s * total + o];
    if (d < bd) { bd = d; bi = part_idx[(size_t)s * total + o]; }   // strict: earlier split (lower index) wins ties
  }
  idx_out[o] = bi;
  if (d2_out != nullptr) d2_out[o] = bd;
}

template <typename T> struct NNCfg;
template <> struct NNCfg<float> { static constexpr int Q = 4; };
template <> struct NNCfg<double> { static constexpr int Q = 2; };

static void nn_plan(int N, int n, int m, int Q, int& gx, int& nsplit, int& tiles_per_split) {
  const int m_pad = nn_m_pad(m);
  gx = (n + NN_THREADS * Q - 1) / (NN_THREADS * Q);
  const int tiles_total = (m_pad + NN_TILE - 1) / NN_TILE;
  const long long base = (long long)gx * N;
  const long long want = 148LL * 8;          // CTAs needed to fill the machine at 8 CTAs / SM
  long long s = (want + base - 1) / base;
  if (s > tiles_total) s = tiles_total;
  if (s > 64) s = 64;
  if (s < 1) s = 1;
  tiles_per_split = (int)((tiles_total + s - 1) / s);
  nsplit = (tiles_total + tiles_per_split - 1) / tiles_per_split;
}

size_t nn_search_ws_bytes(int N, int n, int m, int elem_size) {
  int gx, nsplit, tps;
  nn_plan(N, n, m, elem_size == 4 ? NNCfg<float>::Q : NNCfg<double>::Q, gx, nsplit, tps);
  if (nsplit == 1) return 0;
  return align_up((size_t)nsplit * N * n * elem_size) + align_up((size_t)nsplit * N * n * 4);
}

template <typename T>
int nn_search(const T* q, const double* state, const typename Vec4<T>::type* packed, int N, int n, int m,
              int32_t* idx_out, T* d2_out, const int* stop_flag, void* ws, size_t ws_bytes, cudaStream_t st) {
  constexpr int Q = NNCfg<T>::Q;
  int gx, nsplit, tps;
  nn_plan(N, n, m, Q, gx, nsplit, tps);
  const int m_pad = nn_m_pad(m);
  T* part_d2 = nullptr;
  int32_t* part_idx = nullptr;
  if (nsplit > 1) {
    Arena ar(ws, ws_bytes);
    part_d2 = ar.take<T>((size_t)nsplit * N * n);
    part_idx = ar.take<int32_t>((size_t)nsplit * N * n);
    if (part_d2 == nullptr || part_idx == nullptr) { set_error("nn_search: workspace too small"); return DICP_ERR_WORKSPACE; }
  }
  if (N > 65535) { set_error("nn_search: N > 65535 pairs per call is not supported"); return DICP_ERR_ARG; }
  dim3 grid(gx, N, nsplit);
  nn_scan_kernel<T, Q><<<grid, NN_THREADS, 0, st>>>(q, state, packed, n, m_pad, tps, idx_out, d2_out, part_d2, part_idx, stop_flag);
  DICP_LAUNCH_CHECK();
  if (nsplit > 1) {
    const size_t total = (size_t)N * n;
    nn_merge_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(part_d2, part_idx, total, nsplit, idx_out, d2_out, stop_flag);
    DICP_LAUNCH_CHECK();
  }
  return 0;
}
template int nn_search<float>(const float*, const double*, const float4*, int, int, int, int32_t*, float*, const int*, void*, size_t, cudaStream_t);
template int nn_search<double>(const double*, const double*, const double4*, int, int, int, int32_t*, double*, const int*, void*, size_t, cudaStream_t);

size_t nn_workspace_bytes(int N, int n, int m, int elem_size) {
  return align_up((size_t)N * nn_m_pad(m) * 4 * elem_size) + nn_search_ws_bytes(N, n, m, elem_size);
}

template <typename T>
static int nn_argmin_impl(const T* q, const T* tgt, int N, int n, int m, int S, int32_t* idx_out, T* d2_out,
                          void* ws, size_t ws_bytes, void* stream) {
  DICP_CHECK_ARG(q && tgt && idx_out, "null pointer");
  DICP_CHECK_ARG(N >= 1 && n >= 1 && m >= 1, "N, n, m must be >= 1");
  DICP_CHECK_ARG(S == 3 || S == 6, "S must be 3 or 6");
  if (ws_bytes < nn_workspace_bytes(N, n, m, sizeof(T)) || ws == nullptr) { set_error("dicp_nn_argmin: workspace too small"); return DICP_ERR_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  Arena ar(ws, ws_bytes);
  typedef typename Vec4<T>::type V4;
  V4* packed = ar.take<V4>((size_t)N * nn_m_pad(m));
  int rc = nn_pack_targets<T>(tgt, N, m, S, packed, st);
  if (rc) return rc;
  return nn_search<T>(q, nullptr, packed, N, n, m, idx_out, d2_out, nullptr, (char*)ws + ar.off, ws_bytes - ar.off, st);
}

// ---- gather / scatter-add (nn.py:37-38 and its autograd adjoint) -----------------------
template <typename T>
__global__ void gather_rows_kernel(const T* __restrict__ tgt, const int32_t* __restrict__ idx, int n, int m, int S, T* __restrict__ out) {
  const int b = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * S) return;
  const int i = e / S, c = e - i * S;
  out[((size_t)b * n + i) * S + c] = tgt[((size_t)b * m + idx[(size_t)b * n + i]) * S + c];
}
template <typename T>
__global__ void gather_rows_bwd_kernel(const T* __restrict__ g, const int32_t* __restrict__ idx, int n, int m, int S, T* __restrict__ gt) {
  const int b = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * S) return;
  const int i = e / S, c = e - i * S;
  atomicAdd(&gt[((size_t)b * m + idx[(size_t)b * n + i]) * S + c], g[((size_t)b * n + i) * S + c]);
}

template <typename T>
static int gather_impl(const T* tgt, const int32_t* idx, int N, int n, int m, int S, T* out, void* stream, bool bwd) {
  DICP_CHECK_ARG(tgt && idx && out, "null pointer");
  DICP_CHECK_ARG(N >= 1 && n >= 1 && m >= 1 && (S == 3 || S == 6) && N <= 65535, "bad sizes");
  dim3 grid((n * S + 255) / 256, N);
  if (!bwd) gather_rows_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(tgt, idx, n, m, S, out);
  else gather_rows_bwd_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(tgt, idx, n, m, S, out);
  DICP_LAUNCH_CHECK();
  return 0;
}

}  // namespace dicp

extern "C" {
size_t dicp_nn_workspace_bytes(int N, int n, int m, int elem_size) { return dicp::nn_workspace_bytes(N, n, m, elem_size); }
int dicp_nn_argmin_f32(const float* q, const float* tgt, int N, int n, int m, int S, int32_t* idx_out, float* d2_out,
                       void* ws, size_t ws_bytes, void* stream) {
  return dicp::nn_argmin_impl<float>(q, tgt, N, n, m, S, idx_out, d2_out, ws, ws_bytes, stream);
}
int dicp_nn_argmin_f64(const double* q, const double* tgt, int N, int n, int m, int S, int32_t* idx_out, double* d2_out,
                       void* ws, size_t ws_bytes, void* stream) {
  return dicp::nn_argmin_impl<double>(q, tgt, N, n, m, S, idx_out, d2_out, ws, ws_bytes, stream);
}
size_t dicp_nn_packed_bytes(int N, int m, int elem_size) { return dicp::align_up((size_t)N * dicp::nn_m_pad(m) * 4 * elem_size); }
int dicp_nn_pack_f32(const float* tgt, int N, int m, int S, void* packed, void* stream) {
  DICP_CHECK_ARG(tgt && packed && N >= 1 && N <= 65535 && m >= 1 && (S == 3 || S == 6), "bad arguments");
  return dicp::nn_pack_targets<float>(tgt, N, m, S, (float4*)packed, (cudaStream_t)stream);
}
int dicp_nn_pack_f64(const double* tgt, int N, int m, int S, void* packed, void* stream) {
  DICP_CHECK_ARG(tgt && packed && N >= 1 && N <= 65535 && m >= 1 && (S == 3 || S == 6), "bad arguments");
  return dicp::nn_pack_targets<double>(tgt, N, m, S, (double4*)packed, (cudaStream_t)stream);
}
int dicp_nn_scan_f32(const float* q, const void* packed, int N, int n, int m, int32_t* idx_out, float* d2_out, void* ws,
                     size_t ws_bytes, void* stream) {
  DICP_CHECK_ARG(q && packed && idx_out && N >= 1 && n >= 1 && m >= 1, "bad arguments");
  return dicp::nn_search<float>(q, nullptr, (const float4*)packed, N, n, m, idx_out, d2_out, nullptr, ws, ws_bytes, (cudaStream_t)stream);
}
int dicp_nn_scan_f64(const double* q, const void* packed, int N, int n, int m, int32_t* idx_out, double* d2_out, void* ws,
                     size_t ws_bytes, void* stream) {
  DICP_CHECK_ARG(q && packed && idx_out && N >= 1 && n >= 1 && m >= 1, "bad arguments");
  return dicp::nn_search<double>(q, nullptr, (const double4*)packed, N, n, m, idx_out, d2_out, nullptr, ws, ws_bytes, (cudaStream_t)stream);
}
int dicp_gather_rows_f32(const float* tgt, const int32_t* idx, int N, int n, int m, int S, float* out, void* stream) {
  return dicp::gather_impl<float>(tgt, idx, N, n, m, S, out, stream, false);
}
int dicp_gather_rows_f64(const double* tgt, const int32_t* idx, int N, int n, int m, int S, double* out, void* stream) {
  return dicp::gather_impl<double>(tgt, idx, N, n, m, S, out, stream, false);
}
int dicp_gather_rows_bwd_f32(const float* g, const int32_t* idx, int N, int n, int m, int S, float* gt, void* stream) {
  return dicp::gather_impl<float>(g, idx, N, n, m, S, gt, stream, true);
}
int dicp_gather_rows_bwd_f64(const double* g, const int32_t* idx, int N, int n, int m, int S, double* gt, void* stream) {
  return dicp::gather_impl<double>(g, idx, N, n, m, S, gt, stream, true);
}
}
