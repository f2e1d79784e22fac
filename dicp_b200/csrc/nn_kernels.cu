// nn_kernels.cu — batched brute-force nearest neighbour (replaces cdist+argmin, nn.py:32-35).
//
// Layout: the targets of every pair are repacked once per call into 32-byte records (16 B fp32, 32 B fp64), padded
// with far-away sentinels to a multiple of NN_CHUNK.  A CTA owns THREADS*Q queries of one
// pair (Q queries live in the registers of each thread) and streams the pair's records
// through shared memory in double-buffered cp.async tiles; every thread reads each record
// with one broadcast LDS.128 and evaluates Q distances against it.  This is FP32 CUDA-core
// work (a K=3 distance is not a contraction worth tensor cores).
//
// fp32 (nn_scan_f32_kernel): a filter-and-refine scan in a per-pair translated frame (x' = x - c,
// y' = y - c, c = centre of the targets' bounding box, so that magnitudes are those of the cloud and
// not of its distance from the origin).  Records hold (-2y', |y'|^2), so
//   v_j = |y'_j|^2 - 2 x'.y'_j = d_j^2 - |x'|^2     costs three FFMAs (half the FMA-pipe work of the
// direct form (x-y)^2, and the record operands are served from the operand-reuse cache, which
// is what actually limits FP32 issue on this part: one register read per bank per cycle).
// v_j is not exact (cancellation), so it is only a filter: per chunk of NN_CHUNK records a
// running minimum (FMNMX3, ALU pipe), per query the best chunk, its value and the best value
// among all *other* chunks.  If that runner-up is further away than a rigorous rounding bound
// (nn_filter_margin), the nearest neighbour must be inside the best chunk and is found there
// with the exact direct-form distance on the ORIGINAL coordinates; otherwise the query is appended
// to its pair's list together with the threshold (best filter value + margin) below which a chunk can
// still hold the nearest neighbour, and nn_fallback_filter_kernel re-runs the filter for the listed
// queries and evaluates exactly every chunk under the threshold.  Either way the result is the
// direct-form fp32 argmin with torch.argmin's first-minimum tie-break.  On the benchmark clouds
// (extent +-20 m) about 0.3 % of the queries take the fallback, wherever the clouds sit (without the
// translation a cloud 1 km from the origin sent every query to the fallback: 10x slower).
// fp64 (nn_scan_kernel<double,.>): plain direct-form scan, used by the fp64 parity tests.
//
// The target range is split over blockIdx.z when that is needed to fill 148 SMs or to keep the
// tail wave short; partial (d2, idx) results are merged in split order (earlier split wins ties).
#include "common.cuh"
#include "tile_stream.cuh"

namespace dicp {

template <typename T> struct Sentinel;
template <> struct Sentinel<float> { static __host__ __device__ float far() { return 1.0e18f; } static __device__ float inf() { return __int_as_float(0x7f800000); } };
template <> struct Sentinel<double> { static __host__ __device__ double far() { return 1.0e150; } static __device__ double inf() { return __longlong_as_double(0x7ff0000000000000LL); } };

template <typename T> __device__ __forceinline__ T t_min(T a, T b);
template <> __device__ __forceinline__ float t_min<float>(float a, float b) { return fminf(a, b); }
template <> __device__ __forceinline__ double t_min<double>(double a, double b) { return fmin(a, b); }

// fp32 packed layout per call: [N, m_pad] filter records | [N, m_pad] original (x,y,z,0) records | [N] centres.
// The filter works in a per-pair translated frame (centre c of the targets' bounding box): with
// y' = fl(y - c) the record is (-2y', fl(|y'|^2)) and v = |y'|^2 - 2 x'.y' = |x' - y'|^2 - |x'|^2 costs three
// FFMAs.  Translating keeps |x'| at the size of the cloud instead of its distance from the origin, which is
// what the cancellation error of the expanded form scales with (a cloud 1 km away would otherwise send every
// query to the exact fallback).  Refine and fallback use the original records, so the result is unaffected.
__device__ __forceinline__ void make_record(double x, double y, double z, double4& v) { v.x = x; v.y = y; v.z = z; v.w = 0.0; }

template <typename T>
__global__ void nn_center_kernel(const T* __restrict__ tgt, int m_all, int S, const int32_t* __restrict__ m_valid, float4* __restrict__ centers) {
  __shared__ float smin[3][8], smax[3][8];
  const int b = blockIdx.x;
  const T* tb = tgt + (size_t)b * m_all * S;
  const int m = m_valid != nullptr ? min(max(m_valid[b], 1), m_all) : m_all;
  float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  for (int j = threadIdx.x; j < m; j += blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = (float)tb[(size_t)j * S + c];
      lo[c] = fminf(lo[c], v);
      hi[c] = fmaxf(hi[c], v);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[c] = fminf(lo[c], __shfl_down_sync(0xffffffffu, lo[c], o));
      hi[c] = fmaxf(hi[c], __shfl_down_sync(0xffffffffu, hi[c], o));
    }
    if (lane == 0) { smin[c][warp] = lo[c]; smax[c][warp] = hi[c]; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float cc[3];
    for (int c = 0; c < 3; ++c) {
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { smin[c][0] = fminf(smin[c][0], smin[c][w]); smax[c][0] = fmaxf(smax[c][0], smax[c][w]); }
      cc[c] = 0.5f * smin[c][0] + 0.5f * smax[c][0];
      if (!(fabsf(cc[c]) < 1.0e30f)) cc[c] = 0.0f;          // NaN / inf coordinates: no translation
    }
    centers[b] = make_float4(cc[0], cc[1], cc[2], 0.0f);
  }
}

// NNF_PACKED = 1 (experiment, off): the filter on packed fma.rn.f32x2 (FFMA2), two queries per instruction, explicit
// 64-bit register pairs for the query components and the running sums.  The idea: a 64-bit operand reads both
// register banks at once, so the scalar loop's same-parity collisions between the query component and the running
// sum (29 % of its FFMAs) cannot happen, and the FFMAs need half the issue slots.  The filter records are then stored
// with every component duplicated, (ax,ax,ay,ay | az,az,aw,aw) = 32 bytes, so that a record arrives as four
// ready-made 64-bit operands with two LDS.128.  Measured at C3 (round 2): 10.4 ms per scan against 8.64 ms scalar --
// three 64-bit operands per FFMA2 are more register bandwidth than the file delivers in two cycles -- so the scalar
// loop stays; results are bit-identical either way (same IEEE FMAs).
#ifndef NNF_PACKED
#define NNF_PACKED 0
#endif
constexpr int NNF_FREC = NNF_PACKED ? 2 : 1;        // float4 units per filter record
// views into the fp32 packed buffer: [N, m_pad] filter records | [N, m_pad] original (x,y,z,0) | [N] centres
__host__ __device__ inline const float4* nn_orig_records(const float4* packed, int N, int m_pad) { return packed + (size_t)NNF_FREC * N * m_pad; }
__host__ __device__ inline const float4* nn_centers(const float4* packed, int N, int m_pad) { return packed + (size_t)(NNF_FREC + 1) * N * m_pad; }
size_t nn_packed_bytes(int N, int m, int elem_size) {
  return elem_size == 4 ? align_up(((size_t)N * nn_m_pad(m) * (NNF_FREC + 1) + (size_t)N) * 16) : align_up((size_t)N * nn_m_pad(m) * 32);
}

__global__ void nn_pack_f32_kernel(const float* __restrict__ tgt, int m, int m_pad, int S, const int32_t* __restrict__ m_valid,
                                   const float4* __restrict__ centers, float4* __restrict__ filt, float4* __restrict__ orig) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m_pad) return;
  const int mb = m_valid != nullptr ? min(max(m_valid[b], 1), m) : m;      // rows >= m_valid[b] are padding: never a neighbour
  float4 f, o;
  if (j < mb) {
    const float* row = tgt + ((size_t)b * m + j) * S;
    const float4 c = centers[b];
    const float y0 = (float)((double)row[0] - (double)c.x), y1 = (float)((double)row[1] - (double)c.y),
                y2 = (float)((double)row[2] - (double)c.z);
    f = make_float4(-2.0f * y0, -2.0f * y1, -2.0f * y2, (float)((double)y0 * y0 + (double)y1 * y1 + (double)y2 * y2));
    o = make_float4(row[0], row[1], row[2], 0.0f);
  } else {
    const float far = Sentinel<float>::far();
    f = make_float4(-2.0f * far, -2.0f * far, -2.0f * far, 3.0f * far * far);
    o = make_float4(far, far, far, 0.0f);
  }
  if (NNF_PACKED) {
    filt[((size_t)b * m_pad + j) * 2] = make_float4(f.x, f.x, f.y, f.y);
    filt[((size_t)b * m_pad + j) * 2 + 1] = make_float4(f.z, f.z, f.w, f.w);
  } else {
    filt[(size_t)b * m_pad + j] = f;
  }
  orig[(size_t)b * m_pad + j] = o;
}

__global__ void nn_pack_f64_kernel(const double* __restrict__ tgt, int m, int m_pad, int S, const int32_t* __restrict__ m_valid,
                                   double4* __restrict__ packed) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m_pad) return;
  double4 v;
  const int mb = m_valid != nullptr ? min(max(m_valid[b], 1), m) : m;
  if (j < mb) {
    const double* row = tgt + ((size_t)b * m + j) * S;
    make_record(row[0], row[1], row[2], v);
  } else {
    make_record(Sentinel<double>::far(), Sentinel<double>::far(), Sentinel<double>::far(), v);
  }
  packed[(size_t)b * m_pad + j] = v;
}

template <>
int nn_pack_targets<float>(const float* tgt, int N, int m, int S, const int32_t* m_valid, float4* packed, cudaStream_t st) {
  const int m_pad = nn_m_pad(m);
  float4* orig = packed + (size_t)NNF_FREC * N * m_pad;
  float4* centers = packed + (size_t)(NNF_FREC + 1) * N * m_pad;
  nn_center_kernel<float><<<N, 256, 0, st>>>(tgt, m, S, m_valid, centers);
  DICP_LAUNCH_CHECK();
  dim3 grid((m_pad + 255) / 256, N);
  nn_pack_f32_kernel<<<grid, 256, 0, st>>>(tgt, m, m_pad, S, m_valid, centers, packed, orig);
  DICP_LAUNCH_CHECK();
  return 0;
}
template <>
int nn_pack_targets<double>(const double* tgt, int N, int m, int S, const int32_t* m_valid, double4* packed, cudaStream_t st) {
  const int m_pad = nn_m_pad(m);
  dim3 grid((m_pad + 255) / 256, N);
  nn_pack_f64_kernel<<<grid, 256, 0, st>>>(tgt, m, m_pad, S, m_valid, packed);
  DICP_LAUNCH_CHECK();
  return 0;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

constexpr int NN_THREADS = 128;   // fp64 scan
constexpr int NN_TILE = 512;      // records per shared-memory stage

// Exact direct-form scan over (x,y,z,.) records: every query of the pair (the fp64 path).
template <typename T, int Q, int THREADS>
__global__ void __launch_bounds__(THREADS)
nn_scan_kernel(const T* __restrict__ q, const double* __restrict__ state,
               const typename Vec4<T>::type* __restrict__ packed, int n, int m_pad, int tiles_per_split,
               int32_t* __restrict__ idx_out, T* __restrict__ d2_out, T* __restrict__ part_d2,
               int32_t* __restrict__ part_idx, const int* __restrict__ stop_flag, const int32_t* __restrict__ n_valid) {
  constexpr int NN_THREADS = THREADS;
  typedef typename Vec4<T>::type V4;
  if (stop_flag != nullptr && *stop_flag != 0) return;
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  // rows >= n_valid[b] of a ragged batch are padding: they get index 0 and cost nothing (a CTA of padding exits at once)
  const int nb = n_valid != nullptr ? min(max(n_valid[b], 1), n) : n;
  if (blockIdx.x * (NN_THREADS * Q) >= nb) {
#pragma unroll
    for (int k = 0; k < Q; ++k) {
      const int i = blockIdx.x * (NN_THREADS * Q) + tid + k * NN_THREADS;
      if (i >= n) continue;
      const size_t o = (size_t)b * n + i;
      if (part_d2 == nullptr) { idx_out[o] = 0; if (d2_out != nullptr) d2_out[o] = T(0); }
      else { part_d2[((size_t)blockIdx.z * gridDim.y) * n + o] = Sentinel<T>::inf(); part_idx[((size_t)blockIdx.z * gridDim.y) * n + o] = 0; }
    }
    return;
  }
  __shared__ __align__(16) V4 tile[2][NN_TILE];
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
    if (i >= nb) i = nb - 1;
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
    const int start = t * NN_TILE;
    __syncthreads();
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
    if (i >= nb) {                              // padding row inside a partly valid block
      const size_t o = (size_t)b * n + i;
      if (part_d2 == nullptr) { idx_out[o] = 0; if (d2_out != nullptr) d2_out[o] = T(0); }
      else { part_d2[((size_t)split * N) * n + o] = Sentinel<T>::inf(); part_idx[((size_t)split * N) * n + o] = 0; }
      continue;
    }
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

// ---------------------------------------------------------------------------------------
// fp32: expanded-form filter scan + exact refine + exact fallback
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float min3(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

// Tuning knobs of the filter scan (profiles/r01_nn_tuning.md has the sweeps): records per shared-memory tile, CTAs per SM
// the register budget allows, records interleaved per thread, loop nest of the FFMAs (0 component-major, 1 record-major,
// 2 query-major), and whether / how the query components are pinned to the lanes of a register quad (NNF_XQUAD, NNF_XROT).
constexpr int NNF_THREADS = 128;
#ifndef NNF_TILE_RECORDS
#define NNF_TILE_RECORDS 1024
#endif
#ifndef NNF_MINB
#define NNF_MINB 5
#endif
#ifndef NNF_IL
#define NNF_IL 8
#endif
#ifndef NNF_PIL
#define NNF_PIL 4          // records interleaved per thread in the packed loop
#endif
#ifndef NNF_ORDER
#define NNF_ORDER 1
#endif
#ifndef NNF_XQUAD
#define NNF_XQUAD 1
#endif
#define NNF_FMA(xv, av, cv) __fmaf_rn(xv, av, cv)
constexpr int NNF_TILE = NNF_TILE_RECORDS / NNF_FREC;     // records per shared-memory stage (16 KB; two stages, 5 CTAs per SM)
#if NNF_PACKED
typedef unsigned long long nnf_u64;
__device__ __forceinline__ nnf_u64 ffma2(nnf_u64 a, nnf_u64 b, nnf_u64 c) {
  nnf_u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ nnf_u64 pack2(float lo, float hi) {
  nnf_u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(nnf_u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
#endif

// Filter margin of one query, in the translated frame (x = x', y = y' below).  u = 2^-24.
// The filter value of target y is
//   v = fl(x0*a0 + fl(x1*a1 + fl(x2*a2 + h))),  a = -2y exactly, h = fl(|y|^2) (rounded once from fp64),
// so |v - (d^2 - |x|^2)| <= u (|y|^2 + |s1| + |s2| + |s3|) (1 + 4u) with the exact partial sums
//   s1 = |y|^2 - 2 x2 y2,  s2 = s1 - 2 x1 y1,  s3 = s2 - 2 x0 y0.
// Only targets within rho = 2 d_best of the query can beat the best one, and for those each |s_i|
// differs from its value at y = x (|x|^2 - 2 x2^2, ... , -|x|^2) by at most (2|x| + 2(|x_2|+..)) rho + rho^2.
// A target in another chunk can only be the true (direct-form fp32) nearest neighbour if its filter
// value is within  E_other + E_best + 10 u d_best^2  of the best chunk's.  The translation itself rounds:
// x' = fl(x - c) and y' = fl(y - c) are each off by at most u|x'|, u|y'|, which moves a squared distance of
// a target within rho by at most 2 rho u (2|x'| + rho).  The constants below are inflated for the rounding
// of this bound itself.
__device__ __forceinline__ float nn_filter_margin(float x0, float x1, float x2, float v_best) {
  const float U = 5.9604645e-8f;                                   // 2^-24
  const float xx = x0 * x0 + x1 * x1 + x2 * x2;
  const float xn = sqrtf(xx);
  const float d2 = fmaxf(v_best + xx, 0.0f);
  const float rho = 2.0f * sqrtf(d2) + 1.0e-3f * xn + 1.0e-30f;
  const float s1 = fabsf(xx - 2.0f * x2 * x2);
  const float s2 = fabsf(xx - 2.0f * x2 * x2 - 2.0f * x1 * x1);
  const float E = U * ((xn + rho) * (xn + rho) + s1 + s2 + xx + 14.5f * xn * rho + 3.0f * rho * rho);
  return 2.1f * E + 16.0f * U * d2 + 4.2f * U * rho * (2.0f * xn + rho);
}

template <int Q>
__global__ void __launch_bounds__(NNF_THREADS, NNF_MINB)
nn_scan_f32_kernel(const float* __restrict__ q, const double* __restrict__ state, const float4* __restrict__ packed, int n,
                   int m_pad, int tiles_per_split, int32_t* __restrict__ idx_out, float* __restrict__ d2_out,
                   float* __restrict__ part_d2, int32_t* __restrict__ part_idx, unsigned int* __restrict__ fb_count,
                   int32_t* __restrict__ fb_list, float* __restrict__ fb_thr, int fb_cap, unsigned long long* __restrict__ fb_key,
                   const int* __restrict__ stop_flag, const int32_t* __restrict__ n_valid) {
  if (stop_flag != nullptr && *stop_flag != 0) return;
  __shared__ __align__(16) float4 tile[2][NNF_TILE * NNF_FREC];
  const float INF = Sentinel<float>::inf();
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int split = blockIdx.z;
  const int N = gridDim.y;
  // rows >= n_valid[b] of a ragged batch are padding: they get index 0 and cost nothing (a CTA of padding exits at once)
  const int nb = n_valid != nullptr ? min(max(n_valid[b], 1), n) : n;
  if (blockIdx.x * (NNF_THREADS * Q) >= nb) {
#pragma unroll
    for (int k = 0; k < Q; ++k) {
      const int i = blockIdx.x * (NNF_THREADS * Q) + tid + k * NNF_THREADS;
      if (i >= n) continue;
      const size_t o = (size_t)b * n + i;
      if (part_d2 == nullptr) { idx_out[o] = 0; if (d2_out != nullptr) d2_out[o] = 0.0f; }
      else { part_d2[((size_t)split * N) * n + o] = INF; part_idx[((size_t)split * N) * n + o] = 0; }
    }
    return;
  }
  const float4* tg = packed + (size_t)b * m_pad * NNF_FREC;
  const float4* tg_orig = nn_orig_records(packed, N, m_pad) + (size_t)b * m_pad;
  const float4 ctr = nn_centers(packed, N, m_pad)[b];

  // the query in the target frame (ICP.py:137).  The transform is re-read and the point recomputed
  // identically for the exact refine, so that neither stays in registers during the scan.
  auto query_point = [&](int i, float* xo) {
    const float* p = q + ((size_t)b * n + i) * 3;
    const float pv[3] = {p[0], p[1], p[2]};
    if (state != nullptr) {
      const double* rec = state + (size_t)b * ST_STRIDE;
      float C[9], a[3];
#pragma unroll
      for (int c = 0; c < 9; ++c) C[c] = (float)rec[ST_C + c];
      rotate(C, pv, a);
      xo[0] = a[0] + (float)rec[ST_R]; xo[1] = a[1] + (float)rec[ST_R + 1]; xo[2] = a[2] + (float)rec[ST_R + 2];
    } else {
      xo[0] = pv[0]; xo[1] = pv[1]; xo[2] = pv[2];
    }
  };
  float x[Q][3], best[Q], second[Q];           // x: the query in the translated filter frame
  int bchunk[Q];
  const int q0 = blockIdx.x * (NNF_THREADS * Q) + tid;
#pragma unroll
  for (int k = 0; k < Q; ++k) {
    int i = q0 + k * NNF_THREADS;
    if (i >= nb) i = nb - 1;
    float xo[3];
    query_point(i, xo);
    x[k][0] = xo[0] - ctr.x; x[k][1] = xo[1] - ctr.y; x[k][2] = xo[2] - ctr.z;
    best[k] = INF; second[k] = INF; bchunk[k] = 0;
  }
#if NNF_XQUAD
  // Round trip of (x', best) through shared memory with a 128-bit load: the load's destination is an aligned
  // register quad, which pins x'_0, x'_1, x'_2 of every query to register banks 0, 1, 2 for the whole scan.
  static_assert(NNF_TILE * NNF_FREC >= NNF_THREADS * Q, "the second tile buffer stages the queries");
#pragma unroll
#ifndef NNF_XROT
#define NNF_XROT 3
#endif
  // lane order inside the quad: component c of a record sits in bank c (records also arrive by LDS.128), so the
  // query components are rotated by NNF_XROT lanes to keep x'_c out of the bank of the record component it multiplies
  for (int k = 0; k < Q; ++k) {
    float l[4];
    l[(0 + NNF_XROT) & 3] = x[k][0]; l[(1 + NNF_XROT) & 3] = x[k][1]; l[(2 + NNF_XROT) & 3] = x[k][2]; l[(3 + NNF_XROT) & 3] = best[k];
    tile[1][k * NNF_THREADS + tid] = make_float4(l[0], l[1], l[2], l[3]);
  }
#pragma unroll
  for (int k = 0; k < Q; ++k) {
    const unsigned int addr = (unsigned int)__cvta_generic_to_shared(&tile[1][k * NNF_THREADS + tid]);
    float l[4];
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(l[0]), "=f"(l[1]), "=f"(l[2]), "=f"(l[3]) : "r"(addr));
    x[k][0] = l[(0 + NNF_XROT) & 3]; x[k][1] = l[(1 + NNF_XROT) & 3]; x[k][2] = l[(2 + NNF_XROT) & 3]; best[k] = l[(3 + NNF_XROT) & 3];
  }
  __syncthreads();                               // before the first cp.async into tile[1]
#endif

  const int tiles_total = (m_pad + NNF_TILE - 1) / NNF_TILE;
  const int t_begin = split * tiles_per_split;
  const int t_end = min(t_begin + tiles_per_split, tiles_total);

  auto issue_tile = [&](int t, int buf) {
    const int start = t * NNF_TILE;
    const int cnt = min(NNF_TILE, m_pad - start) * NNF_FREC;
    for (int u = tid; u < cnt; u += NNF_THREADS) cp_async16(&tile[buf][u], tg + (size_t)start * NNF_FREC + u);
    cp_async_commit();
  };
#if NNF_PACKED
  // query pairs (2 kp, 2 kp + 1) as 64-bit operands, formed once
  nnf_u64 x2[Q / 2][3];
#pragma unroll
  for (int kp = 0; kp < Q / 2; ++kp)
#pragma unroll
    for (int c = 0; c < 3; ++c) x2[kp][c] = pack2(x[2 * kp][c], x[2 * kp + 1][c]);
#endif

  if (t_begin < t_end) issue_tile(t_begin, 0);
  for (int t = t_begin; t < t_end; ++t) {
    const int buf = (t - t_begin) & 1;
    if (t + 1 < t_end) { issue_tile(t + 1, buf ^ 1); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncthreads();
    const int start = t * NNF_TILE;
    const int nchunks = min(NNF_TILE, m_pad - start) / NN_CHUNK;
    for (int c = 0; c < nchunks; ++c) {
      float cm[Q];
#pragma unroll
      for (int k = 0; k < Q; ++k) cm[k] = INF;
      const float4* ch = &tile[buf][c * NN_CHUNK * NNF_FREC];
#if NNF_PACKED
#pragma unroll
      for (int j = 0; j < NN_CHUNK; j += NNF_PIL) {
        // two broadcast LDS.128 per record deliver its four duplicated components as 64-bit operands
        nnf_u64 axx[NNF_PIL], ayy[NNF_PIL], azz[NNF_PIL], aww[NNF_PIL];
        nnf_u64 v2[NNF_PIL][Q / 2];
#pragma unroll
        for (int i = 0; i < NNF_PIL; ++i) {
          const unsigned int addr = (unsigned int)__cvta_generic_to_shared(ch + (j + i) * 2);
          asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(axx[i]), "=l"(ayy[i]) : "r"(addr));
          asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(azz[i]), "=l"(aww[i]) : "r"(addr + 16));
        }
#pragma unroll
        for (int i = 0; i < NNF_PIL; ++i) {
#pragma unroll
          for (int kp = 0; kp < Q / 2; ++kp) v2[i][kp] = ffma2(x2[kp][2], azz[i], aww[i]);
#pragma unroll
          for (int kp = 0; kp < Q / 2; ++kp) v2[i][kp] = ffma2(x2[kp][1], ayy[i], v2[i][kp]);
#pragma unroll
          for (int kp = 0; kp < Q / 2; ++kp) v2[i][kp] = ffma2(x2[kp][0], axx[i], v2[i][kp]);
        }
#pragma unroll
        for (int i = 0; i < NNF_PIL; i += 2)
#pragma unroll
          for (int kp = 0; kp < Q / 2; ++kp) {
            float a0, a1, b0, b1;
            unpack2(v2[i][kp], a0, a1);
            unpack2(v2[i + 1][kp], b0, b1);
            cm[2 * kp] = min3(cm[2 * kp], a0, b0);
            cm[2 * kp + 1] = min3(cm[2 * kp + 1], a1, b1);
          }
      }
#else
#pragma unroll
      for (int j = 0; j < NN_CHUNK; j += NNF_IL) {
        // one broadcast LDS.128 per record; the record's components sit in the same operand slot of
        // Q consecutive FFMAs, so they are served from the operand-reuse cache, not the register file
        float4 a[NNF_IL];
        float v[NNF_IL][Q];
#pragma unroll
        for (int i = 0; i < NNF_IL; ++i) a[i] = ch[j + i];
#if NNF_ORDER == 0
#pragma unroll
        for (int i = 0; i < NNF_IL; ++i)
#pragma unroll
          for (int k = 0; k < Q; ++k) v[i][k] = NNF_FMA(x[k][2], a[i].z, a[i].w);
#pragma unroll
        for (int i = 0; i < NNF_IL; ++i)
#pragma unroll
          for (int k = 0; k < Q; ++k) v[i][k] = NNF_FMA(x[k][1], a[i].y, v[i][k]);
#pragma unroll
        for (int i = 0; i < NNF_IL; ++i)
#pragma unroll
          for (int k = 0; k < Q; ++k) v[i][k] = NNF_FMA(x[k][0], a[i].x, v[i][k]);
#elif NNF_ORDER == 1
#pragma unroll
        for (int i = 0; i < NNF_IL; ++i) {
#pragma unroll
          for (int k = 0; k < Q; ++k) v[i][k] = NNF_FMA(x[k][2], a[i].z, a[i].w);
#pragma unroll
          for (int k = 0; k < Q; ++k) v[i][k] = NNF_FMA(x[k][1], a[i].y, v[i][k]);
#pragma unroll
          for (int k = 0; k < Q; ++k) v[i][k] = NNF_FMA(x[k][0], a[i].x, v[i][k]);
        }
#else
#pragma unroll
        for (int k = 0; k < Q; ++k) {
#pragma unroll
          for (int i = 0; i < NNF_IL; ++i) v[i][k] = NNF_FMA(x[k][2], a[i].z, a[i].w);
#pragma unroll
          for (int i = 0; i < NNF_IL; ++i) v[i][k] = NNF_FMA(x[k][1], a[i].y, v[i][k]);
#pragma unroll
          for (int i = 0; i < NNF_IL; ++i) v[i][k] = NNF_FMA(x[k][0], a[i].x, v[i][k]);
        }
#endif
#pragma unroll
        for (int i = 0; i < NNF_IL; i += 2)
#pragma unroll
          for (int k = 0; k < Q; ++k) cm[k] = min3(cm[k], v[i][k], v[i + 1][k]);
      }
#endif
#pragma unroll
      for (int k = 0; k < Q; ++k) {
        second[k] = fminf(second[k], fmaxf(cm[k], best[k]));     // smallest minimum among the chunks that are not the best
        if (cm[k] < best[k]) { best[k] = cm[k]; bchunk[k] = start / NN_CHUNK + c; }
      }
    }
    __syncthreads();
  }

  // refine: exact direct-form argmin inside the winning chunk; queries whose runner-up chunk is
  // within the error margin go to the exact fallback scan
#pragma unroll
  for (int k = 0; k < Q; ++k) {
    const int i = q0 + k * NNF_THREADS;
    if (i >= n) continue;
    const size_t o = (size_t)b * n + i;
    int found = 0;
    float bd = INF;
    if (i >= nb) {                              // padding row inside a partly valid block
      if (part_d2 == nullptr) { idx_out[o] = 0; if (d2_out != nullptr) d2_out[o] = 0.0f; }
      else { part_d2[((size_t)split * N) * n + o] = INF; part_idx[((size_t)split * N) * n + o] = 0; }
      continue;
    }
    if (t_begin < t_end) {
      const float margin = nn_filter_margin(x[k][0], x[k][1], x[k][2], best[k]);
      if (!(second[k] - best[k] > margin)) {
        const unsigned int slot = atomicAdd(&fb_count[b], 1u);
        if (slot < (unsigned int)fb_cap) {
          fb_list[(size_t)b * fb_cap + slot] = i;
          // the true nearest neighbour's filter value is at most best + margin: the fallback only has to
          // look into chunks whose filter minimum is below this threshold (slack for rounding the sum)
          const float thr = best[k] + margin;
          fb_thr[(size_t)b * fb_cap + slot] = thr + 4.8e-7f * fabsf(thr);
        }
        fb_key[o] = 0xffffffffffffffffull;
      }
      float xo[3];
      query_point(i, xo);
      const float4* ch = tg_orig + (size_t)bchunk[k] * NN_CHUNK;
#pragma unroll 4
      for (int j = 0; j < NN_CHUNK; ++j) {
        const float4 y = ch[j];
        const float d = dist2<float>(xo[0], xo[1], xo[2], y.x, y.y, y.z);
        if (d < bd) { bd = d; found = bchunk[k] * NN_CHUNK + j; }
      }
    }
    if (part_d2 == nullptr) {
      idx_out[o] = found;
      if (d2_out != nullptr) d2_out[o] = bd;
    } else {
      part_d2[((size_t)split * N) * n + o] = bd;
      part_idx[((size_t)split * N) * n + o] = found;
    }
  }
}

// Exact fallback for the filter's ambiguous queries (fp32).  The scan kernel lists each of them with the
// threshold  thr = (its best filter value) + (its error margin):  the direct-form nearest neighbour can only
// lie in a 32-record chunk whose filter minimum is <= thr (nn_filter_margin).  One warp per CTA walks a pair's
// list: it runs the same three-FFMA filter over its share of the target range, queues every (query, chunk)
// candidate in shared memory, and evaluates the queued chunks with the exact direct-form distance on the
// ORIGINAL records, one candidate per lane.  Results merge across chunks, CTAs and target splits through a
// 64-bit atomicMin on (d2 bits, index): the lexicographic first minimum, as torch.argmin.  Typically 2-3
// chunks per query are candidates, so this costs half of an exact rescan of all targets.
constexpr int FBF_THREADS = 32;
#ifndef FBF_QPL
#define FBF_QPL 2
#endif
constexpr int FBF_Q = FBF_QPL;          // listed queries per lane and pass
constexpr int FBF_SLOT_BITS = 8;        // queue entries: (chunk << 8) | query slot
#ifndef FBF_TILE_RECORDS
#define FBF_TILE_RECORDS 512
#endif
constexpr int FBF_TILE = FBF_TILE_RECORDS / NNF_FREC;
constexpr int FBF_QUEUE = 512;

__global__ void __launch_bounds__(FBF_THREADS)
nn_fallback_filter_kernel(const float* __restrict__ q, const double* __restrict__ state, const float4* __restrict__ packed, int n, int m_pad,
                          int tiles_per_split, const unsigned int* __restrict__ qcount, const int32_t* __restrict__ qlist,
                          const float* __restrict__ qthr, int qcap, unsigned long long* __restrict__ qkey,
                          const int* __restrict__ stop_flag) {
  if (stop_flag != nullptr && *stop_flag != 0) return;
  __shared__ __align__(16) float4 tile[2][FBF_TILE * NNF_FREC];
  __shared__ float qx[FBF_THREADS * FBF_Q][3];
  __shared__ int qid[FBF_THREADS * FBF_Q];
  __shared__ int queue[FBF_QUEUE];
  const int lane = threadIdx.x;
  const int b = blockIdx.y;
  const int split = blockIdx.z;
  const int N = gridDim.y;
  const int nq = (int)min(qcount[b], (unsigned int)qcap);
  const float4* tg = packed + (size_t)b * m_pad * NNF_FREC;
  const float4* tg_orig = nn_orig_records(packed, N, m_pad) + (size_t)b * m_pad;
  const float4 ctr = nn_centers(packed, N, m_pad)[b];
  const float INF = Sentinel<float>::inf();
  const int tiles_total = (m_pad + FBF_TILE - 1) / FBF_TILE;
  const int t_begin = split * tiles_per_split;
  const int t_end = min(t_begin + tiles_per_split, tiles_total);
  if (t_begin >= t_end) return;

  // evaluates the queued (query, chunk) candidates exactly, one per lane
  auto flush = [&](int count) {
    __syncwarp();
    for (int t = lane; t < count; t += FBF_THREADS) {
      const int e = queue[t];
      const int slot = e & ((1 << FBF_SLOT_BITS) - 1), chunk = e >> FBF_SLOT_BITS;
      const float x0 = qx[slot][0], x1 = qx[slot][1], x2 = qx[slot][2];
      const float4* ch = tg_orig + (size_t)chunk * NN_CHUNK;
      float bd = INF;
      int found = chunk * NN_CHUNK;
#pragma unroll 4
      for (int j = 0; j < NN_CHUNK; ++j) {
        const float4 y = ch[j];
        const float d = dist2<float>(x0, x1, x2, y.x, y.y, y.z);
        if (d < bd) { bd = d; found = chunk * NN_CHUNK + j; }
      }
      atomicMin(&qkey[(size_t)b * n + qid[slot]], ((unsigned long long)__float_as_uint(bd) << 32) | (unsigned int)found);
    }
    __syncwarp();
  };
  static_assert(FBF_THREADS * FBF_Q <= (1 << FBF_SLOT_BITS), "queue entries keep the query slot in FBF_SLOT_BITS bits");

  for (int qblk = blockIdx.x * (FBF_THREADS * FBF_Q); qblk < nq; qblk += gridDim.x * (FBF_THREADS * FBF_Q)) {
    float x[FBF_Q][3], thr[FBF_Q];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < FBF_Q; ++k) {
      const int s = qblk + k * FBF_THREADS + lane;
      const bool valid = s < nq;
      const int sc = valid ? s : nq - 1;
      const int i = qlist[(size_t)b * qcap + sc];
      thr[k] = valid ? qthr[(size_t)b * qcap + sc] : -INF;          // -inf: never a candidate
      // the query in the target frame, with the arithmetic of the scan kernel (ICP.py:137)
      const float* p = q + ((size_t)b * n + i) * 3;
      const float pv[3] = {p[0], p[1], p[2]};
      float xo[3];
      if (state != nullptr) {
        const double* rec = state + (size_t)b * ST_STRIDE;
        float C[9], a[3];
#pragma unroll
        for (int c = 0; c < 9; ++c) C[c] = (float)rec[ST_C + c];
        rotate(C, pv, a);
        xo[0] = a[0] + (float)rec[ST_R]; xo[1] = a[1] + (float)rec[ST_R + 1]; xo[2] = a[2] + (float)rec[ST_R + 2];
      } else {
        xo[0] = pv[0]; xo[1] = pv[1]; xo[2] = pv[2];
      }
      x[k][0] = xo[0] - ctr.x; x[k][1] = xo[1] - ctr.y; x[k][2] = xo[2] - ctr.z;
      const int slot = k * FBF_THREADS + lane;
      qx[slot][0] = xo[0]; qx[slot][1] = xo[1]; qx[slot][2] = xo[2];
      qid[slot] = i;
    }
    __syncwarp();

    auto issue_tile = [&](int t, int buf) {
      const int start = t * FBF_TILE;
      const int cnt = min(FBF_TILE, m_pad - start) * NNF_FREC;
      for (int u = lane; u < cnt; u += FBF_THREADS) cp_async16(&tile[buf][u], tg + (size_t)start * NNF_FREC + u);
      cp_async_commit();
    };
    int queued = 0;                          // warp-uniform
    issue_tile(t_begin, 0);
    for (int t = t_begin; t < t_end; ++t) {
      const int buf = (t - t_begin) & 1;
      if (t + 1 < t_end) { issue_tile(t + 1, buf ^ 1); cp_async_wait<1>(); }
      else cp_async_wait<0>();
      __syncwarp();
      const int start = t * FBF_TILE;
      const int nchunks = min(FBF_TILE, m_pad - start) / NN_CHUNK;
      for (int c = 0; c < nchunks; ++c) {
        float cm[FBF_Q];
#pragma unroll
        for (int k = 0; k < FBF_Q; ++k) cm[k] = INF;
        const float4* ch = &tile[buf][c * NN_CHUNK * NNF_FREC];
#pragma unroll
        for (int j = 0; j < NN_CHUNK; j += 2) {
#if NNF_PACKED
          const float4 p0 = ch[2 * j], p1 = ch[2 * j + 1], p2 = ch[2 * j + 2], p3 = ch[2 * j + 3];
          const float4 a0 = make_float4(p0.x, p0.z, p1.x, p1.z), a1 = make_float4(p2.x, p2.z, p3.x, p3.z);
#else
          const float4 a0 = ch[j], a1 = ch[j + 1];
#endif
#pragma unroll
          for (int k = 0; k < FBF_Q; ++k) {
            float v0 = __fmaf_rn(x[k][2], a0.z, a0.w), v1 = __fmaf_rn(x[k][2], a1.z, a1.w);
            v0 = __fmaf_rn(x[k][1], a0.y, v0); v1 = __fmaf_rn(x[k][1], a1.y, v1);
            v0 = __fmaf_rn(x[k][0], a0.x, v0); v1 = __fmaf_rn(x[k][0], a1.x, v1);
            cm[k] = min3(cm[k], v0, v1);
          }
        }
        const int chunk = start / NN_CHUNK + c;
#pragma unroll
        for (int k = 0; k < FBF_Q; ++k) {
          const bool hit = !(cm[k] > thr[k]);
          const unsigned int mask = __ballot_sync(0xffffffffu, hit);
          if (mask != 0u) {
            if (hit) queue[queued + __popc(mask & ((1u << lane) - 1u))] = (chunk << FBF_SLOT_BITS) | (k * FBF_THREADS + lane);
            queued += __popc(mask);
            if (queued > FBF_QUEUE - FBF_THREADS) { flush(queued); queued = 0; }
          }
        }
      }
      __syncwarp();                         // everybody is done with tile[buf] before it is refilled
    }
    flush(queued);
  }
}

// writes the fallback's (d2, index) keys of the listed queries to the outputs
__global__ void nn_fallback_unpack_kernel(const unsigned int* __restrict__ qcount, const int32_t* __restrict__ qlist, int qcap, int n,
                                          const unsigned long long* __restrict__ qkey, int32_t* __restrict__ idx_out,
                                          float* __restrict__ d2_out, const int* __restrict__ stop_flag) {
  if (stop_flag != nullptr && *stop_flag != 0) return;
  const int b = blockIdx.y;
  const int cnt = (int)min(qcount[b], (unsigned int)qcap);
  for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < cnt; slot += gridDim.x * blockDim.x) {
    const size_t o = (size_t)b * n + qlist[(size_t)b * qcap + slot];
    const unsigned long long key = qkey[o];
    idx_out[o] = (int32_t)(key & 0xffffffffull);
    if (d2_out != nullptr) d2_out[o] = __uint_as_float((unsigned int)(key >> 32));
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

// grid: x = query blocks, y = pair, z = target split.  Split when the query grid alone cannot
// fill the machine, and when it gives only a few waves (a short last wave costs up to one
// CTA duration: keep CTAs short relative to the whole launch).
struct NNPlan { int q, gx, nsplit, tiles_per_split, threads, tile; };

static NNPlan nn_plan(int N, int n, int m, int elem_size) {
  NNPlan p;
  const int m_pad = nn_m_pad(m);
  int ctas_per_sm;
  if (elem_size == 4) {
    // 8 queries per thread, 93 registers, 5 CTAs / SM (profiles/r01_nn_tuning.md).  The split heuristic below
    // counts 4 per SM: at C3 (4096 CTAs) a fifth resident CTA must not tip the launch into two target splits
    p.threads = NNF_THREADS; p.tile = NNF_TILE; p.q = 8;
    ctas_per_sm = 4;
  } else {
    p.threads = NN_THREADS; p.tile = NN_TILE; p.q = 2; ctas_per_sm = 4;
  }
  p.gx = (n + p.threads * p.q - 1) / (p.threads * p.q);
  const int tiles_total = (m_pad + p.tile - 1) / p.tile;
  const long long base = (long long)p.gx * N;
  const long long slots = (long long)device_sm_count() * ctas_per_sm;
#ifndef NN_WAVES
#define NN_WAVES 6
#endif
#ifndef NN_MIN_TILES
#define NN_MIN_TILES 8
#endif
  const long long want = slots * NN_WAVES;   // >= 6 waves => the short last wave costs a few percent at most
  long long s = (want + base - 1) / base;
  const long long max_by_work = tiles_total / NN_MIN_TILES > 0 ? tiles_total / NN_MIN_TILES : 1;   // >= 8 tiles per CTA keeps prologue/refine overhead small
  const long long fill = (slots + base - 1) / base;                            // but always fill the machine
  if (s > max_by_work) s = max_by_work > fill ? max_by_work : fill;
  if (s > tiles_total) s = tiles_total;
  if (s > 64) s = 64;
  if (s < 1) s = 1;
  p.tiles_per_split = (int)((tiles_total + s - 1) / s);
  p.nsplit = (tiles_total + p.tiles_per_split - 1) / p.tiles_per_split;
  return p;
}

// workspace of one search: [split partials d2 | split partials idx] [fallback counter] [fallback list]
struct NNWorkspace { void* part_d2; int32_t* part_idx; unsigned int* fb_count; int32_t* fb_list; float* fb_thr; int fb_cap; unsigned long long* fb_key; size_t total; };
static NNWorkspace nn_ws_layout(void* ws, size_t ws_bytes, int N, int n, int m, int elem_size) {
  const NNPlan p = nn_plan(N, n, m, elem_size);
  Arena ar(ws, ws_bytes);
  NNWorkspace w;
  w.part_d2 = nullptr; w.part_idx = nullptr; w.fb_count = nullptr; w.fb_list = nullptr; w.fb_thr = nullptr; w.fb_key = nullptr;
  if (p.nsplit > 1) {
    w.part_d2 = ar.take<char>((size_t)p.nsplit * N * n * elem_size);
    w.part_idx = ar.take<int32_t>((size_t)p.nsplit * N * n);
  }
  w.fb_cap = 0;
  if (elem_size == 4) {
    w.fb_cap = n * p.nsplit;                                       // every (query, split) can flag at most once
    w.fb_count = ar.take<unsigned int>(N);                         // per pair
    w.fb_list = ar.take<int32_t>((size_t)N * w.fb_cap);            // per pair: query indices
    w.fb_thr = ar.take<float>((size_t)N * w.fb_cap);               // per listed query: filter threshold of its candidate chunks
    w.fb_key = ar.take<unsigned long long>((size_t)N * n);         // per query: (d2 bits << 32 | index) of the exact rescan
  }
  w.total = ar.off;
  return w;
}

size_t nn_search_ws_bytes(int N, int n, int m, int elem_size) { return nn_ws_layout(nullptr, 0, N, n, m, elem_size).total; }

static int launch_scan(const NNPlan& p, const NNWorkspace& w, int N, cudaStream_t st, const float* q, const double* state, const float4* packed,
                       int n, int m_pad, int32_t* idx_out, float* d2_out, const int* stop_flag, const int32_t* n_valid) {
  dim3 grid(p.gx, N, p.nsplit);
  DICP_CUDA(cudaMemsetAsync(w.fb_count, 0, sizeof(unsigned int) * N, st));
  nn_scan_f32_kernel<8><<<grid, NNF_THREADS, 0, st>>>(q, state, packed, n, m_pad, p.tiles_per_split, idx_out, d2_out, (float*)w.part_d2,
                                                         w.part_idx, w.fb_count, w.fb_list, w.fb_thr, w.fb_cap, w.fb_key, stop_flag, n_valid);
  DICP_LAUNCH_CHECK();
  return 0;
}
static int launch_scan(const NNPlan& p, const NNWorkspace& w, int N, cudaStream_t st, const double* q, const double* state, const double4* packed,
                       int n, int m_pad, int32_t* idx_out, double* d2_out, const int* stop_flag, const int32_t* n_valid) {
  dim3 grid(p.gx, N, p.nsplit);
  nn_scan_kernel<double, 2, NN_THREADS><<<grid, NN_THREADS, 0, st>>>(q, state, packed, n, m_pad, p.tiles_per_split, idx_out, d2_out,
                                                                     (double*)w.part_d2, w.part_idx, stop_flag, n_valid);
  DICP_LAUNCH_CHECK();
  return 0;
}
// exact fallback for the fp32 filter's ambiguous queries (nn_fallback_filter_kernel): one warp per CTA (a pair
// rarely lists more than a few dozen queries), the target range split over blockIdx.z so that the whole
// machine works on it; candidates merge through atomicMin on fb_key.
#ifndef NN_FB_TILES_PER_SPLIT
#define NN_FB_TILES_PER_SPLIT 4
#endif
#ifndef NN_FB_MAX_SPLITS
#define NN_FB_MAX_SPLITS 8
#endif
static int launch_fallback(const NNWorkspace& w, int N, cudaStream_t st, const float* q, const double* state, const float4* packed, int n, int m_pad,
                           int32_t* idx_out, float* d2_out, const int* stop_flag) {
  const int tiles_total = (m_pad + FBF_TILE - 1) / FBF_TILE;
  int fs = tiles_total / NN_FB_TILES_PER_SPLIT;
  if (fs > NN_FB_MAX_SPLITS) fs = NN_FB_MAX_SPLITS;
  if (fs < 1) fs = 1;
  const int tps = (tiles_total + fs - 1) / fs;
  fs = (tiles_total + tps - 1) / tps;
  // CTAs stride over the lists (their lengths are only known on the device).  Few CTAs per pair when
  // there are many pairs (empty CTAs are not free), enough to fill the machine when there are few.
  const int gx_cap = (w.fb_cap + FBF_THREADS * FBF_Q - 1) / (FBF_THREADS * FBF_Q);
  int gx = (device_sm_count() * 8 + N * fs - 1) / (N * fs);
#ifndef NN_FB_MIN_GX
#define NN_FB_MIN_GX 4
#endif
  if (gx < NN_FB_MIN_GX) gx = NN_FB_MIN_GX;
  if (gx > gx_cap) gx = gx_cap;
  dim3 grid(gx, N, fs);
  nn_fallback_filter_kernel<<<grid, FBF_THREADS, 0, st>>>(q, state, packed, n, m_pad, tps, w.fb_count, w.fb_list, w.fb_thr, w.fb_cap, w.fb_key, stop_flag);
  DICP_LAUNCH_CHECK();
  dim3 ugrid(1, N);
  nn_fallback_unpack_kernel<<<ugrid, 256, 0, st>>>(w.fb_count, w.fb_list, w.fb_cap, n, w.fb_key, idx_out, d2_out, stop_flag);
  DICP_LAUNCH_CHECK();
  return 0;
}
static int launch_fallback(const NNWorkspace&, int, cudaStream_t, const double*, const double*, const double4*, int, int, int32_t*, double*, const int*) {
  return 0;   // the fp64 scan is exact (direct form)
}

template <typename T>
int nn_search(const T* q, const double* state, const typename Vec4<T>::type* packed, int N, int n, int m,
              int32_t* idx_out, T* d2_out, const int* stop_flag, void* ws, size_t ws_bytes, cudaStream_t st, const int32_t* n_valid) {
  if (N > 65535) { set_error("nn_search: N > 65535 pairs per call is not supported"); return DICP_ERR_ARG; }
  const NNPlan p = nn_plan(N, n, m, sizeof(T));
  const int m_pad = nn_m_pad(m);
  const NNWorkspace w = nn_ws_layout(ws, ws_bytes, N, n, m, sizeof(T));
  if (w.total > 0 && (ws == nullptr || w.total > ws_bytes)) { set_error("nn_search: workspace too small"); return DICP_ERR_WORKSPACE; }
  int rc = launch_scan(p, w, N, st, q, state, packed, n, m_pad, idx_out, d2_out, stop_flag, n_valid);
  if (rc) return rc;
  if (p.nsplit > 1) {
    const size_t total = (size_t)N * n;
    nn_merge_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const T*)w.part_d2, w.part_idx, total, p.nsplit, idx_out, d2_out, stop_flag);
    DICP_LAUNCH_CHECK();
  }
  return launch_fallback(w, N, st, q, state, packed, n, m_pad, idx_out, d2_out, stop_flag);
}
template int nn_search<float>(const float*, const double*, const float4*, int, int, int, int32_t*, float*, const int*, void*, size_t, cudaStream_t, const int32_t*);
template int nn_search<double>(const double*, const double*, const double4*, int, int, int, int32_t*, double*, const int*, void*, size_t, cudaStream_t, const int32_t*);

size_t nn_workspace_bytes(int N, int n, int m, int elem_size) {
  return nn_packed_bytes(N, m, elem_size) + nn_search_ws_bytes(N, n, m, elem_size);
}

template <typename T>
static int nn_argmin_impl(const T* q, const T* tgt, int N, int n, int m, int S, const int32_t* m_valid, int32_t* idx_out, T* d2_out,
                          void* ws, size_t ws_bytes, void* stream) {
  DICP_CHECK_ARG(q && tgt && idx_out, "null pointer");
  DICP_CHECK_ARG(N >= 1 && n >= 1 && m >= 1, "N, n, m must be >= 1");
  DICP_CHECK_ARG(S == 3 || S == 6, "S must be 3 or 6");
  if (ws_bytes < nn_workspace_bytes(N, n, m, sizeof(T)) || ws == nullptr) { set_error("dicp_nn_argmin: workspace too small"); return DICP_ERR_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  Arena ar(ws, ws_bytes);
  typedef typename Vec4<T>::type V4;
  V4* packed = (V4*)ar.take<char>(nn_packed_bytes(N, m, sizeof(T)));
  int rc = nn_pack_targets<T>(tgt, N, m, S, m_valid, packed, st);
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

// ---- one pair, targets split over GPUs: merge of the slice results over NVLink peer memory -----
// Every rank folds the (distance, global index) key of each query into the key buffer of EVERY rank
// with a 64-bit system-scope atomic MIN (red.min.u64 over NVLink); after a cross-GPU barrier each
// buffer holds the global nearest neighbour.  The bit pattern of a non-negative fp32 is monotonic
// in its value, so integer MIN == nearest neighbour with the first-minimum tie-break (nn.py:35-36).
constexpr int NN_MAX_PEERS = 16;
struct PeerKeys { unsigned long long* p[NN_MAX_PEERS]; };

__global__ void __launch_bounds__(256)
nn_push_keys_kernel(const float* __restrict__ d2, const int32_t* __restrict__ idx, int offset, size_t count, PeerKeys peers, int world,
                    int first_peer, unsigned long long* __restrict__ reset_buf) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned long long key = ((unsigned long long)__float_as_uint(d2[i]) << 32) | (unsigned int)(idx[i] + offset);
    for (int j = 0; j < world; ++j) {
      int r = first_peer + j;                  // every rank starts at a different peer: spreads the traffic over the links
      if (r >= world) r -= world;
      atomicMin_system(peers.p[r] + i, key);
    }
    if (reset_buf != nullptr) reset_buf[i] = ~0ull;      // the buffer of the iteration after this one
  }
}

__global__ void __launch_bounds__(256)
nn_unpack_keys_kernel(const unsigned long long* __restrict__ keys, size_t count, int32_t* __restrict__ idx_out, float* __restrict__ d2_out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned long long key = keys[i];
    idx_out[i] = (int32_t)(key & 0xffffffffull);
    if (d2_out != nullptr) d2_out[i] = __uint_as_float((unsigned int)(key >> 32));
  }
}

// FP32 CUDA-core issue-rate probe: what a loop of independent FFMAs reaches on this device (the denominator of the
// scan's roofline; MEASURED_PEAKS.json has no FP32 figure).  16 independent accumulator chains per thread, 8 warps per
// CTA, 8 CTAs per SM; operands live in registers only.  bench.py times it with CUDA events.
__global__ void __launch_bounds__(256)
probe_ffma_kernel(int iters, float seed, float* __restrict__ sink) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + (float)(threadIdx.x + i);
  const float m = 1.0f + seed * 1e-9f, c = seed * 1e-7f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = __fmaf_rn(a[i], m, c);
    }
  }
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 123.456f) sink[0] = s;               // keeps the chains alive, practically never true
}

static int grid_for(size_t count) {
  const size_t b = (count + 255) / 256;
  const size_t cap = (size_t)device_sm_count() * 8;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace dicp

extern "C" {
size_t dicp_nn_workspace_bytes(int N, int n, int m, int elem_size) { return dicp::nn_workspace_bytes(N, n, m, elem_size); }
int dicp_nn_argmin_f32(const float* q, const float* tgt, int N, int n, int m, int S, const int32_t* m_valid, int32_t* idx_out,
                       float* d2_out, void* ws, size_t ws_bytes, void* stream) {
  return dicp::nn_argmin_impl<float>(q, tgt, N, n, m, S, m_valid, idx_out, d2_out, ws, ws_bytes, stream);
}
int dicp_nn_argmin_f64(const double* q, const double* tgt, int N, int n, int m, int S, const int32_t* m_valid, int32_t* idx_out,
                       double* d2_out, void* ws, size_t ws_bytes, void* stream) {
  return dicp::nn_argmin_impl<double>(q, tgt, N, n, m, S, m_valid, idx_out, d2_out, ws, ws_bytes, stream);
}
size_t dicp_nn_packed_bytes(int N, int m, int elem_size) { return dicp::nn_packed_bytes(N, m, elem_size); }
int dicp_nn_pack_f32(const float* tgt, int N, int m, int S, const int32_t* m_valid, void* packed, void* stream) {
  DICP_CHECK_ARG(tgt && packed && N >= 1 && N <= 65535 && m >= 1 && (S == 3 || S == 6), "bad arguments");
  return dicp::nn_pack_targets<float>(tgt, N, m, S, m_valid, (float4*)packed, (cudaStream_t)stream);
}
int dicp_nn_pack_f64(const double* tgt, int N, int m, int S, const int32_t* m_valid, void* packed, void* stream) {
  DICP_CHECK_ARG(tgt && packed && N >= 1 && N <= 65535 && m >= 1 && (S == 3 || S == 6), "bad arguments");
  return dicp::nn_pack_targets<double>(tgt, N, m, S, m_valid, (double4*)packed, (cudaStream_t)stream);
}
int dicp_nn_scan_f32(const float* q, const double* state, const void* packed, int N, int n, int m, int32_t* idx_out, float* d2_out,
                     void* ws, size_t ws_bytes, void* stream) {
  DICP_CHECK_ARG(q && packed && idx_out && N >= 1 && n >= 1 && m >= 1, "bad arguments");
  return dicp::nn_search<float>(q, state, (const float4*)packed, N, n, m, idx_out, d2_out, nullptr, ws, ws_bytes, (cudaStream_t)stream);
}
int dicp_nn_scan_f64(const double* q, const double* state, const void* packed, int N, int n, int m, int32_t* idx_out, double* d2_out,
                     void* ws, size_t ws_bytes, void* stream) {
  DICP_CHECK_ARG(q && packed && idx_out && N >= 1 && n >= 1 && m >= 1, "bad arguments");
  return dicp::nn_search<double>(q, state, (const double4*)packed, N, n, m, idx_out, d2_out, nullptr, ws, ws_bytes, (cudaStream_t)stream);
}
int dicp_nn_push_keys_f32(const float* d2, const int32_t* idx, int offset, size_t count, const uint64_t* peer_bufs, int world,
                          int first_peer, void* reset_buf, void* stream) {
  DICP_CHECK_ARG(d2 && idx && peer_bufs && count >= 1 && world >= 1 && world <= dicp::NN_MAX_PEERS && first_peer >= 0 && first_peer < world,
                 "bad arguments");
  dicp::PeerKeys pk;
  for (int r = 0; r < dicp::NN_MAX_PEERS; ++r) pk.p[r] = r < world ? (unsigned long long*)(uintptr_t)peer_bufs[r] : nullptr;
  dicp::nn_push_keys_kernel<<<dicp::grid_for(count), 256, 0, (cudaStream_t)stream>>>(d2, idx, offset, count, pk, world, first_peer,
                                                                                     (unsigned long long*)reset_buf);
  DICP_LAUNCH_CHECK();
  return 0;
}
int dicp_nn_unpack_keys_f32(const void* keys, size_t count, int32_t* idx_out, float* d2_out, void* stream) {
  DICP_CHECK_ARG(keys && idx_out && count >= 1, "bad arguments");
  dicp::nn_unpack_keys_kernel<<<dicp::grid_for(count), 256, 0, (cudaStream_t)stream>>>((const unsigned long long*)keys, count, idx_out, d2_out);
  DICP_LAUNCH_CHECK();
  return 0;
}
double dicp_probe_fp32_flops(int iters, float* sink, void* stream) {
  if (iters < 1 || sink == nullptr) return 0.0;
  const int grid = dicp::device_sm_count() * 8;
  dicp::probe_ffma_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(iters, 1.0f, sink);
  if (cudaGetLastError() != cudaSuccess) return 0.0;
  return 2.0 * 16.0 * 8.0 * (double)iters * 256.0 * (double)grid;      // FLOP of the launch
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
