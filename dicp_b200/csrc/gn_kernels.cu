// gn_kernels.cu — the Gauss-Newton half of the dICP loop and the forward driver.
//
//   gn_accum_kernel   per correspondence: gather the neighbour row, residual, point-to-plane
//                     projection, trim / robust weights, Jacobian row(s), and the 21+6+4
//                     sums of the normal equations (ICP.py:143-200).  HBM-bound: 48 B per
//                     correspondence (pt2pl).  Per-thread FP32/FP64 partial sums, warp
//                     shuffle + shared-memory tree in double, one partial row per CTA.
//   gn_update_kernel  per pair: fixed-order sum of the CTA partials (deterministic), 6x6
//                     (3x3 for dim 2) pivoted solve in double, exp-map update of (C, r),
//                     side outputs, convergence / freeze bookkeeping (ICP.py:200-260).
//   finalize_kernel   pc = C p + r, T, stats of non-converged pairs (ICP.py:267-292).
//
// The iteration loop runs in dicp_icp_forward_*: one launch sequence per iteration, no
// host synchronisation when const_iter is set, otherwise one flag read every
// `check_every` iterations (the reference synchronises every iteration, ICP.py:240).
#include <stdarg.h>

#include "common.cuh"

namespace dicp {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

constexpr int GN_THREADS = 128;
#ifndef GN_PTS
#define GN_PTS 16
#endif
constexpr int GN_PTS_PER_THREAD = GN_PTS;  // 8 pipelined passes of 2: amortises the 31-value block reduction

int bw_blocks_per_pair(int n) {          // backward point pass: 8 correspondences per thread (2 passes of 4)
  int nb = (n + GN_THREADS * 8 - 1) / (GN_THREADS * 8);
  if (nb < 1) nb = 1;
  if (nb > 128) nb = 128;
  return nb;
}

int gn_blocks_per_pair(int n) {
  int nb = (n + GN_THREADS * GN_PTS_PER_THREAD - 1) / (GN_THREADS * GN_PTS_PER_THREAD);
  if (nb < 1) nb = 1;
  if (nb > 128) nb = 128;
  return nb;
}

// block-wide sum of NV per-thread values: warp shuffle tree in the compute type (32 partials of
// a handful of terms each), then double across warps; result lands in out[0..NV) of the CTA's
// partial row.  Fixed tree => deterministic for a given launch shape.
template <typename T, int NV>
__device__ __forceinline__ void block_reduce_store(const T* vals, double* out, double* smem /* [warps][NV] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    T x = vals[v];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (lane == 0) smem[warp * NV + v] = (double)x;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0.0;
    for (int w = 0; w < nwarps; ++w) s += smem[w * NV + threadIdx.x];
    out[threadIdx.x] = s;
  }
}

template <typename T> struct GnUnroll;
#ifndef GN_U
#define GN_U 2
#endif
template <> struct GnUnroll<float> { static constexpr int U = GN_U; };   // correspondences per thread and pipeline stage
template <> struct GnUnroll<double> { static constexpr int U = 1; };

#ifndef GN_MINB
#define GN_MINB 4
#endif

template <typename T, bool PT2PL, int DIM, bool FASTW>
__global__ void __launch_bounds__(GN_THREADS, sizeof(T) == 4 ? GN_MINB : 1)
gn_accum_kernel(const T* __restrict__ source, const T* __restrict__ target, const void* __restrict__ rows,
                const T* __restrict__ w_init, const int32_t* __restrict__ idx, StateView state, int n, int m, int S,
                WeightParams P, T* __restrict__ weights_out, long long w_pair_stride,
                double* __restrict__ partials, const int* __restrict__ stop_flag) {
  if (stop_flag != nullptr && *stop_flag != 0) return;
  __shared__ double red_smem[(GN_THREADS / 32) * RED_N];
  constexpr int U = GnUnroll<T>::U;
  const int b = blockIdx.y;
  const double* rec = state.base + (size_t)b * state.stride;
  T C[9], r[3];
#pragma unroll
  for (int i = 0; i < 9; ++i) C[i] = (T)rec[state.c + i];
#pragma unroll
  for (int i = 0; i < 3; ++i) r[i] = (T)rec[state.r + i];
  const T active = state.active >= 0 ? (T)rec[state.active] : T(1);

  T acc[RED_N];
#pragma unroll
  for (int i = 0; i < RED_N; ++i) acc[i] = T(0);

  const T* src_b = source + (size_t)b * n * 3;
  const T* tgt_b = target + (size_t)b * m * S;
  const size_t row_b = (size_t)b * m;
  const bool packed = sizeof(T) == 4 && rows != nullptr;
  const int32_t* idx_b = idx + (size_t)b * n;
  constexpr int R = PT2PL ? 1 : 3;
  T* wout = weights_out != nullptr ? weights_out + (size_t)b * w_pair_stride : nullptr;

  // Three-stage software pipeline in registers, U correspondences per thread and stage: while pass t
  // is evaluated, the row gathers of pass t+1 and the index / point / weight loads of pass t+2 are in
  // flight (the gather address depends on the index, so one stage cannot cover both latencies).
  // Without it a warp alternates between waiting for HBM and computing, and with 16 warps per SM
  // (the 31 accumulators cost the registers) nothing else fills the gaps.
  struct StageA { int ii[U], jj[U]; T p[U][3], w0[U]; };
  struct StageB { T q[U][3], nr[U][3]; };
  auto load_a = [&](StageA& a, int base) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = base + u * GN_THREADS + threadIdx.x;
      a.ii[u] = i < n ? i : -1;
      const int ic = i < n ? i : n - 1;
      a.jj[u] = idx_b[ic];
      a.p[u][0] = src_b[3 * ic]; a.p[u][1] = src_b[3 * ic + 1]; a.p[u][2] = src_b[3 * ic + 2];
      a.w0[u] = (w_init != nullptr ? w_init[(size_t)b * n + ic] : T(1)) * active;          // ICP.py:257
    }
  };
  auto load_b = [&](StageB& g, const StageA& a) {
    if (packed) {
#pragma unroll
      for (int u = 0; u < U; ++u) load_packed_row<T, PT2PL>(rows, row_b + a.jj[u], g.q[u], g.nr[u]);
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u) load_target_row<T, PT2PL>(tgt_b + (size_t)a.jj[u] * S, S, g.q[u], g.nr[u]);
    }
  };
  // one pipeline step: loads of pass t+2 into `an2`, gathers of pass t+1 (an1 -> gn1), arithmetic of pass t (ac, gc)
  auto step = [&](const StageA& ac, const StageB& gc, const StageA& an1, StageB& gn1, StageA& an2, int base, int stride) {
    load_a(an2, base + 2 * stride);
    load_b(gn1, an1);
    // branch-free over the U correspondences so that their long dependent chains (sqrt, tanh, div)
    // interleave; lanes past the end of the cloud evaluate a clamped copy and are masked out
    PointEval<T> s[U];
#pragma unroll
    for (int u = 0; u < U; ++u) eval_point<T, PT2PL, FASTW>(C, r, ac.p[u], gc.q[u], gc.nr[u], ac.w0[u], P, s[u]);
    if (base + GN_THREADS * U <= n) {
#pragma unroll
      for (int u = 0; u < U; ++u) accum_point<T, PT2PL, DIM>(acc, s[u], ac.w0[u], (T)P.thresh);
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (ac.ii[u] >= 0) accum_point<T, PT2PL, DIM>(acc, s[u], ac.w0[u], (T)P.thresh);
    }
    if (wout != nullptr) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (ac.ii[u] < 0) continue;
#pragma unroll
        for (int k = 0; k < R; ++k) wout[(size_t)ac.ii[u] * R + k] = s[u].w;                   // ICP.py:222, 164-165
      }
    }
  };
  const int stride = gridDim.x * (GN_THREADS * U);
  int base = blockIdx.x * (GN_THREADS * U);
  if (base < n) {
    // the three stage buffers rotate by name over a loop unrolled by three: a register-to-register
    // rotation at the end of each pass would wait for the loads it has just issued
    StageA a0, a1, a2;
    StageB g0, g1, g2;
    load_a(a0, base);
    load_a(a1, base + stride);
    load_b(g0, a0);
    for (;;) {
      step(a0, g0, a1, g1, a2, base, stride);
      if ((base += stride) >= n) break;
      step(a1, g1, a2, g2, a0, base, stride);
      if ((base += stride) >= n) break;
      step(a2, g2, a0, g0, a1, base, stride);
      if ((base += stride) >= n) break;
    }
  }
  block_reduce_store<T, RED_N>(acc, partials + ((size_t)b * gridDim.x + blockIdx.x) * RED_N, red_smem);
}

// ---- fp32 gather table ------------------------------------------------------------------
size_t gn_rows_bytes(int N, int m, int elem_size) { return elem_size == 4 ? (size_t)N * m * 32 : 0; }

template <bool PT2PL>
__global__ void __launch_bounds__(256)
gn_pack_rows_kernel(const float* __restrict__ target, size_t total, int S, float4* __restrict__ rows) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const float* t = target + i * S;
    if (PT2PL) {
      const float2 a = *reinterpret_cast<const float2*>(t), b = *reinterpret_cast<const float2*>(t + 2),
                   c = *reinterpret_cast<const float2*>(t + 4);
      rows[2 * i] = make_float4(a.x, a.y, b.x, b.y);
      rows[2 * i + 1] = make_float4(c.x, c.y, 0.f, 0.f);
    } else {
      rows[i] = make_float4(t[0], t[1], t[2], 0.f);
    }
  }
}

template <typename T>
int gn_pack_rows(const T* target, int N, int m, int S, bool pt2pl, void* rows, cudaStream_t st) {
  if constexpr (sizeof(T) == 4) {
    const size_t total = (size_t)N * m;
    const int grid = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
    if (pt2pl) gn_pack_rows_kernel<true><<<grid, 256, 0, st>>>(target, total, S, (float4*)rows);
    else gn_pack_rows_kernel<false><<<grid, 256, 0, st>>>(target, total, S, (float4*)rows);
    DICP_LAUNCH_CHECK();
  }
  return 0;
}
template int gn_pack_rows<float>(const float*, int, int, int, bool, void*, cudaStream_t);
template int gn_pack_rows<double>(const double*, int, int, int, bool, void*, cudaStream_t);

template <typename T>
int gn_accumulate(const T* source, const T* target, const void* rows, const T* w_init, const int32_t* idx, StateView state,
                  int N, int n, int m, int S, const dicp_params* p, T* weights_out, long long w_pair_stride,
                  double* partials, const int* stop_flag, cudaStream_t st) {
  const WeightParams P = make_weight_params(p);
  dim3 grid(gn_blocks_per_pair(n), N);
  const bool pl = p->icp_type == DICP_PT2PL;
  const bool fast = weight_params_fast(P);
#define DICP_GN_LAUNCH(PL, DM)                                                                                                          \
  do {                                                                                                                                  \
    if (fast) gn_accum_kernel<T, PL, DM, true><<<grid, GN_THREADS, 0, st>>>(source, target, rows, w_init, idx, state, n, m, S, P,        \
                                                                            weights_out, w_pair_stride, partials, stop_flag);           \
    else gn_accum_kernel<T, PL, DM, false><<<grid, GN_THREADS, 0, st>>>(source, target, rows, w_init, idx, state, n, m, S, P,            \
                                                                        weights_out, w_pair_stride, partials, stop_flag);               \
  } while (0)
  if (pl && p->dim == 3) DICP_GN_LAUNCH(true, 3);
  else if (pl) DICP_GN_LAUNCH(true, 2);
  else if (p->dim == 3) DICP_GN_LAUNCH(false, 3);
  else DICP_GN_LAUNCH(false, 2);
#undef DICP_GN_LAUNCH
  DICP_LAUNCH_CHECK();
  return 0;
}
template int gn_accumulate<float>(const float*, const float*, const void*, const float*, const int32_t*, StateView, int, int, int, int,
                                  const dicp_params*, float*, long long, double*, const int*, cudaStream_t);
template int gn_accumulate<double>(const double*, const double*, const void*, const double*, const int32_t*, StateView, int, int, int, int,
                                   const dicp_params*, double*, long long, double*, const int*, cudaStream_t);

// sum the CTA partial rows of pair b in a fixed order (thread v sums slot v)
__device__ __forceinline__ void sum_partials(const double* partials, int b, int nblk, int nslots, double* red) {
  if ((int)threadIdx.x < nslots) {
    double s = 0.0;
    for (int j = 0; j < nblk; ++j) s += partials[((size_t)b * nblk + j) * nslots + threadIdx.x];
    red[threadIdx.x] = s;
  }
  __syncthreads();
}

template <typename T>
__global__ void init_state_kernel(const T* __restrict__ T_init, int N, double* __restrict__ state0,
                                  uint8_t* __restrict__ converged, T* __restrict__ num_iters, T* __restrict__ match_ratio) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= N) return;
  const T* Ti = T_init + (size_t)b * 16;
  double* rec = state0 + (size_t)b * ST_STRIDE;
  for (int i = 0; i < ST_STRIDE; ++i) rec[i] = 0.0;
  for (int rr = 0; rr < 3; ++rr) {
    for (int c = 0; c < 3; ++c) rec[ST_C + rr * 3 + c] = (double)Ti[rr * 4 + c];       // ICP.py:125
    rec[ST_R + rr] = (double)Ti[rr * 4 + 3];                                            // ICP.py:126
  }
  rec[ST_ACTIVE] = 1.0;
  converged[b] = 0;
  num_iters[b] = T(0);
  match_ratio[b] = T(0);
}

template <typename T>
__global__ void __launch_bounds__(128)
gn_update_kernel(const double* __restrict__ partials, int nblk, double* __restrict__ state_cur, double* __restrict__ state_nxt,
                 int k, int dim, int const_iter, double tolerance, T* __restrict__ deltas, T* __restrict__ costs,
                 T* __restrict__ weights, int K, long long nw, uint8_t* __restrict__ converged,
                 T* __restrict__ num_iters, T* __restrict__ match_ratio, Ctrl* __restrict__ ctrl, int N) {
  if (ctrl->stop != 0) return;
  __shared__ double red[RED_N];
  __shared__ int s_dead;
  const int b = blockIdx.x;
  sum_partials(partials, b, nblk, RED_N, red);
  if (threadIdx.x == 0) {
    double* cur = state_cur + (size_t)b * ST_STRIDE;
    double* nxt = state_nxt + (size_t)b * ST_STRIDE;
    for (int i = 0; i < ST_STRIDE; ++i) nxt[i] = 0.0;
    double ni = (double)num_iters[b], mr = (double)match_ratio[b];
    const double prev_cost = k > 0 ? (double)costs[(size_t)b * K + k - 1] : 0.0;
    PairOutcome out;
    pair_update<T>(red, cur, nxt, k, dim, const_iter != 0, tolerance, prev_cost, &ni, &mr, out);
#pragma unroll
    for (int i = 0; i < 6; ++i) deltas[((size_t)b * K + k) * 6 + i] = (T)out.delta[i];   // ICP.py:220
    costs[(size_t)b * K + k] = (T)out.cost;                                              // ICP.py:229-234
    num_iters[b] = (T)ni;
    match_ratio[b] = (T)mr;
    if (out.below) converged[b] = 1;                                                     // ICP.py:239
    s_dead = out.dead_weights ? 1 : 0;
    if (out.singular) atomicExch(&ctrl->singular, 1);
    if (out.below) atomicAdd(&ctrl->nbelow, 1);
    __threadfence();
    const int t = atomicAdd(&ctrl->ticket, 1);
    if (t == N - 1) {                       // last pair of this iteration
      const int nb = atomicAdd(&ctrl->nbelow, 0);
      if (!const_iter && nb == N) ctrl->stop = 1;                                        // ICP.py:259-260
      ctrl->iters_run = k + 1;
      ctrl->ticket = 0;
      ctrl->nbelow = 0;
      __threadfence();
    }
  }
  __syncthreads();
  if (s_dead && weights != nullptr) {                                                    // ICP.py:224-226
    T* wk = weights + ((size_t)b * K + k) * nw;
    const T* wp = wk - nw;
    for (long long i = threadIdx.x; i < nw; i += blockDim.x) wk[i] = wp[i];
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
finalize_kernel(const T* __restrict__ source, int n, const double* __restrict__ state_all, int N, const Ctrl* __restrict__ ctrl,
                double thresh, double rows, T* __restrict__ pc, T* __restrict__ T_out,
                T* __restrict__ num_iters, T* __restrict__ match_ratio) {
  const int b = blockIdx.y;
  const int kr = ctrl->iters_run;
  const double* after = state_all + ((size_t)kr * N + b) * ST_STRIDE;
  T C[9], r[3];
#pragma unroll
  for (int i = 0; i < 9; ++i) C[i] = (T)after[ST_C + i];
#pragma unroll
  for (int i = 0; i < 3; ++i) r[i] = (T)after[ST_R + i];
  const T* src_b = source + (size_t)b * n * 3;
  T* pc_b = pc + (size_t)b * n * 3;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const T p[3] = {src_b[3 * i], src_b[3 * i + 1], src_b[3 * i + 2]};
    T a[3];
    rotate(C, p, a);
    pc_b[3 * i] = a[0] + r[0]; pc_b[3 * i + 1] = a[1] + r[1]; pc_b[3 * i + 2] = a[2] + r[2];   // ICP.py:274
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    T* To = T_out + (size_t)b * 16;                                                      // ICP.py:277-280
    for (int rr = 0; rr < 3; ++rr) {
      for (int c = 0; c < 3; ++c) To[rr * 4 + c] = C[rr * 3 + c];
      To[rr * 4 + 3] = r[rr];
    }
    To[12] = T(0); To[13] = T(0); To[14] = T(0); To[15] = T(1);
    if (kr > 0) {
      const double* last = state_all + ((size_t)(kr - 1) * N + b) * ST_STRIDE;
      double ni = (double)num_iters[b], mr = (double)match_ratio[b];
      final_stats(last, after, kr, thresh, rows, &ni, &mr);                              // ICP.py:267-271
      num_iters[b] = (T)ni;
      match_ratio[b] = (T)mr;
    }
  }
}

// ---------------------------------------------------------------------------------------
// workspace layout of one forward / backward call
// ---------------------------------------------------------------------------------------
IcpWorkspace plan_workspace(void* ws, size_t ws_bytes, int N, int n, int m, int elem_size) {
  Arena ar(ws, ws_bytes);
  IcpWorkspace w;
  w.packed = ar.take<char>(nn_packed_bytes(N, m, elem_size));
  w.nn_ws_bytes = nn_search_ws_bytes(N, n, m, elem_size);
  w.nn_ws = ar.take<char>(w.nn_ws_bytes > 0 ? w.nn_ws_bytes : 1);
  w.partials = ar.take<double>((size_t)N * gn_blocks_per_pair(n) * RED_N);
  w.ctrl = ar.take<Ctrl>(1);
  w.bw_state = ar.take<double>((size_t)N * BW_STRIDE);
  w.bw_partials = ar.take<double>((size_t)N * bw_blocks_per_pair(n) * BW_NRED);
  w.grid_ws_bytes = grid_ws_layout(nullptr, 0, N, n, m, elem_size).total;
  w.grid_ws = ar.take<char>(w.grid_ws_bytes);
  w.rows = ar.take<char>(gn_rows_bytes(N, m, elem_size) > 0 ? gn_rows_bytes(N, m, elem_size) : 1);
  w.total = ar.off;
  return w;
}

template <typename T>
static int icp_forward_impl(const T* source, const T* target, const T* T_init, const T* w_init, const int32_t* m_valid, int N, int n,
                            int m, int S, const dicp_params* p, T* pc, T* T_out, T* deltas, T* weights, T* costs, uint8_t* converged,
                            T* num_iters, T* match_ratio, int32_t* saved_idx, double* saved_state, int32_t* iters_run_host,
                            void* ws, size_t ws_bytes, void* stream, dicp_nn_hook hook = nullptr, void* hook_user = nullptr) {
  DICP_CHECK_ARG(source && target && T_init && p && pc && T_out && deltas && costs && converged && num_iters && match_ratio &&
                 saved_idx && saved_state && iters_run_host, "null pointer");
  DICP_CHECK_ARG(N >= 1 && N <= 65535 && n >= 1 && m >= 1, "need 1 <= N <= 65535, n >= 1, m >= 1");
  DICP_CHECK_ARG(p->dim == 2 || p->dim == 3, "dim must be 2 or 3");                      // ICP.py:79
  DICP_CHECK_ARG(p->icp_type == DICP_PT2PL || p->icp_type == DICP_PT2PT, "icp_type must be pt2pl or pt2pt");
  DICP_CHECK_ARG(p->icp_type != DICP_PT2PL || S == 6, "pt2pl needs a 6-column target");  // ICP.py:103
  DICP_CHECK_ARG(S == 3 || S == 6, "target must have 3 or 6 columns");
  DICP_CHECK_ARG(p->loss_id >= 0 && p->loss_id <= 2, "invalid loss id");                 // loss.py:19
  DICP_CHECK_ARG(p->max_iterations >= 1, "max_iterations must be >= 1");
  DICP_CHECK_ARG(hook != nullptr || p->nn_mode == DICP_NN_BRUTE || p->nn_mode == DICP_NN_GRID, "invalid nn_mode");
  if (ws == nullptr || ws_bytes < plan_workspace(nullptr, 0, N, n, m, sizeof(T)).total) {
    set_error("dicp_icp_forward: workspace too small");
    return DICP_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  typedef typename Vec4<T>::type V4;
  IcpWorkspace w = plan_workspace(ws, ws_bytes, N, n, m, sizeof(T));
  const int K = p->max_iterations;
  const bool pl = p->icp_type == DICP_PT2PL;
  const long long nw = (long long)n * (pl ? 1 : 3);
  const int nblk = gn_blocks_per_pair(n);
  const int check_every = p->check_every >= 1 ? p->check_every : 1;
  const size_t rec_stride = (size_t)N * ST_STRIDE;

  DICP_CUDA(cudaMemsetAsync(w.ctrl, 0, sizeof(Ctrl), st));
  init_state_kernel<T><<<(N + 127) / 128, 128, 0, st>>>(T_init, N, saved_state, converged, num_iters, match_ratio);
  DICP_LAUNCH_CHECK();
  int rc = 0;
  // optional spatial index of the (static) targets; a cloud so clustered that one cell holds a
  // large share of it (e.g. the 1000 x max(source) padding rows of a ragged batch stretch the
  // bounding box) would make the ring search degenerate, so such calls keep the tiled scan
  bool use_grid = hook == nullptr && p->nn_mode == DICP_NN_GRID;
  GridWorkspace gw;
  if (use_grid) {
    gw = grid_ws_layout(w.grid_ws, w.grid_ws_bytes, N, n, m, sizeof(T));
    rc = grid_build<T>(target, N, m, S, m_valid, gw, st);
    if (rc) return rc;
    rc = grid_sort_queries<T>(source, saved_state, N, n, gw, st);      // at the initial transform (record 0)
    if (rc) return rc;
    int max_occ = 0;
    DICP_CUDA(cudaMemcpyAsync(&max_occ, gw.max_occupancy, sizeof(int), cudaMemcpyDeviceToHost, st));
    DICP_CUDA(cudaStreamSynchronize(st));
    if (max_occ > 64 && max_occ > m / 64) use_grid = false;
  }
  if (!use_grid && hook == nullptr) {
    rc = nn_pack_targets<T>(target, N, m, S, m_valid, (V4*)w.packed, st);
    if (rc) return rc;
  }

  rc = gn_pack_rows<T>(target, N, m, S, pl, w.rows, st);
  if (rc) return rc;
  const void* rows = sizeof(T) == 4 ? w.rows : nullptr;

  Ctrl h;
  memset(&h, 0, sizeof(h));
  for (int k = 0; k < K; ++k) {
    double* rec_k = saved_state + (size_t)k * rec_stride;
    int32_t* idx_k = saved_idx + (size_t)k * N * n;
    if (hook != nullptr) {
      rc = hook(hook_user, k, rec_k, idx_k);     // the caller's search (e.g. targets split over GPUs), stream-ordered
      if (rc) { set_error("dicp_icp_forward: the correspondence hook failed at iteration %d (code %d)", k, rc); return rc; }
    } else if (use_grid) rc = grid_search<T>(source, rec_k, N, n, m, gw, true, idx_k, nullptr, &w.ctrl->stop, st);
    else rc = nn_search<T>(source, rec_k, (const V4*)w.packed, N, n, m, idx_k, nullptr, &w.ctrl->stop, w.nn_ws, w.nn_ws_bytes, st);
    if (rc) return rc;
    rc = gn_accumulate<T>(source, target, rows, w_init, idx_k, state_view_records(rec_k), N, n, m, S, p,
                          weights != nullptr ? weights + (size_t)k * nw : nullptr, (long long)K * nw, w.partials, &w.ctrl->stop, st);
    if (rc) return rc;
    gn_update_kernel<T><<<N, 128, 0, st>>>(w.partials, nblk, rec_k, rec_k + rec_stride, k, p->dim, p->const_iter, p->tolerance,
                                           deltas, costs, weights, K, nw, converged, num_iters, match_ratio, w.ctrl, N);
    DICP_LAUNCH_CHECK();
    if (!p->const_iter && ((k + 1) % check_every == 0) && k + 1 < K) {
      DICP_CUDA(cudaMemcpyAsync(&h, w.ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, st));
      DICP_CUDA(cudaStreamSynchronize(st));
      if (h.stop) break;
    }
  }
  dim3 fgrid(gn_blocks_per_pair(n), N);
  finalize_kernel<T><<<fgrid, 256, 0, st>>>(source, n, saved_state, N, w.ctrl, p->match_ratio_thresh, (double)nw,
                                            pc, T_out, num_iters, match_ratio);
  DICP_LAUNCH_CHECK();
  DICP_CUDA(cudaMemcpyAsync(&h, w.ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, st));
  DICP_CUDA(cudaStreamSynchronize(st));
  *iters_run_host = h.iters_run;
  if (h.singular) {
    set_error("dicp_icp_forward: singular normal-equation matrix (torch.linalg.inv would raise, ICP.py:201)");
    return DICP_ERR_SINGULAR;
  }
  return 0;
}

// stage API: the solve / update half of one iteration (ICP.py:200-217, 239) on the partial rows of the
// accumulate stage: fixed-order sum, pivoted solve in double, exp-map update.  One thread per pair is enough.
template <typename T>
__global__ void solve_update_stage_kernel(const double* __restrict__ partials, int nblk, const double* __restrict__ C_r, int N, int dim,
                                          double tolerance, T* __restrict__ delta, double* __restrict__ C_r_next,
                                          uint8_t* __restrict__ below, int* __restrict__ singular) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= N) return;
  double red[RED_N];
  for (int v = 0; v < RED_N; ++v) {
    double s = 0.0;
    for (int j = 0; j < nblk; ++j) s += partials[((size_t)b * nblk + j) * RED_N + v];
    red[v] = s;
  }
  double cur[ST_STRIDE], nxt[ST_STRIDE];
  for (int i = 0; i < ST_STRIDE; ++i) { cur[i] = 0.0; nxt[i] = 0.0; }
  for (int i = 0; i < 12; ++i) cur[i] = C_r[(size_t)b * 12 + i];          // ST_C = 0, ST_R = 9
  cur[ST_ACTIVE] = 1.0;
  double ni = 0.0, mr = 0.0;
  PairOutcome out;
  pair_update<T>(red, cur, nxt, 0, dim, true, tolerance, 0.0, &ni, &mr, out);
  for (int i = 0; i < 6; ++i) delta[(size_t)b * 6 + i] = (T)out.delta[i];
  for (int i = 0; i < 12; ++i) C_r_next[(size_t)b * 12 + i] = nxt[i];
  if (below != nullptr) below[b] = out.below ? 1 : 0;
  if (out.singular && singular != nullptr) atomicExch(singular, 1);
}

template <typename T>
static int solve_update_api(const double* partials, int nblk, const double* C_r, int N, const dicp_params* p, T* delta,
                            double* C_r_next, uint8_t* below, void* stream) {
  DICP_CHECK_ARG(partials && C_r && p && delta && C_r_next, "null pointer");
  DICP_CHECK_ARG(N >= 1 && nblk >= 1 && (p->dim == 2 || p->dim == 3), "bad sizes");
  static_assert(ST_C == 0 && ST_R == 9, "the stage API copies (C, r) to the front of a state record");
  solve_update_stage_kernel<T><<<(N + 63) / 64, 64, 0, (cudaStream_t)stream>>>(partials, nblk, C_r, N, p->dim, p->tolerance, delta,
                                                                              C_r_next, below, nullptr);
  DICP_LAUNCH_CHECK();
  return 0;
}

// stage API: the loop's accumulate kernel in ONE launch on a bare [N,12] (C, r); the per-CTA partial
// rows come back as they are (gn_update_kernel sums them in index order in the loop)
template <typename T>
static int gn_accumulate_api(const T* source, const T* target, const void* rows, const T* w_init, const int32_t* idx, const double* C_r,
                             int N, int n, int m, int S, const dicp_params* p, T* weights_out, double* partials_out, void* stream) {
  DICP_CHECK_ARG(source && target && idx && C_r && p && partials_out, "null pointer");
  DICP_CHECK_ARG(N >= 1 && N <= 65535 && n >= 1 && m >= 1 && (S == 3 || S == 6), "bad sizes");
  DICP_CHECK_ARG(p->icp_type != DICP_PT2PL || S == 6, "pt2pl needs a 6-column target");
  const long long nw = (long long)n * (p->icp_type == DICP_PT2PL ? 1 : 3);
  return gn_accumulate<T>(source, target, sizeof(T) == 4 ? rows : nullptr, w_init, idx, StateView{C_r, 12, 0, 9, -1}, N, n, m, S, p,
                          weights_out, nw, partials_out, nullptr, (cudaStream_t)stream);
}

}  // namespace dicp

extern "C" {
int dicp_version(void) { return DICP_VERSION; }
const char* dicp_last_error(void) { return dicp::g_err; }

size_t dicp_icp_workspace_bytes(int N, int n, int m, int elem_size) {
  return dicp::plan_workspace(nullptr, 0, N, n, m, elem_size).total;
}

int dicp_icp_forward_f32(const float* source, const float* target, const float* T_init, const float* w_init, const int32_t* m_valid,
                         int N, int n, int m, int S, const dicp_params* params, float* pc, float* T_out, float* deltas, float* weights, float* costs,
                         uint8_t* converged, float* num_iters, float* match_ratio, int32_t* saved_idx, double* saved_state,
                         int32_t* iters_run_host, void* ws, size_t ws_bytes, void* stream) {
  return dicp::icp_forward_impl<float>(source, target, T_init, w_init, m_valid, N, n, m, S, params, pc, T_out, deltas, weights, costs,
                                       converged, num_iters, match_ratio, saved_idx, saved_state, iters_run_host, ws, ws_bytes, stream);
}
int dicp_icp_forward_f64(const double* source, const double* target, const double* T_init, const double* w_init, const int32_t* m_valid,
                         int N, int n, int m, int S, const dicp_params* params, double* pc, double* T_out, double* deltas, double* weights,
                         double* costs, uint8_t* converged, double* num_iters, double* match_ratio, int32_t* saved_idx,
                         double* saved_state, int32_t* iters_run_host, void* ws, size_t ws_bytes, void* stream) {
  return dicp::icp_forward_impl<double>(source, target, T_init, w_init, m_valid, N, n, m, S, params, pc, T_out, deltas, weights, costs,
                                        converged, num_iters, match_ratio, saved_idx, saved_state, iters_run_host, ws, ws_bytes, stream);
}
int dicp_icp_forward_hooked_f32(const float* source, const float* target, const float* T_init, const float* w_init, const int32_t* m_valid,
                                int N, int n, int m, int S, const dicp_params* params, float* pc, float* T_out, float* deltas, float* weights,
                                float* costs, uint8_t* converged, float* num_iters, float* match_ratio, int32_t* saved_idx,
                                double* saved_state, int32_t* iters_run_host, void* ws, size_t ws_bytes, void* stream, dicp_nn_hook hook,
                                void* user) {
  if (hook == nullptr) { dicp::set_error("dicp_icp_forward_hooked: null hook"); return DICP_ERR_ARG; }
  return dicp::icp_forward_impl<float>(source, target, T_init, w_init, m_valid, N, n, m, S, params, pc, T_out, deltas, weights, costs,
                                       converged, num_iters, match_ratio, saved_idx, saved_state, iters_run_host, ws, ws_bytes, stream, hook, user);
}
int dicp_icp_forward_hooked_f64(const double* source, const double* target, const double* T_init, const double* w_init, const int32_t* m_valid,
                                int N, int n, int m, int S, const dicp_params* params, double* pc, double* T_out, double* deltas, double* weights,
                                double* costs, uint8_t* converged, double* num_iters, double* match_ratio, int32_t* saved_idx,
                                double* saved_state, int32_t* iters_run_host, void* ws, size_t ws_bytes, void* stream, dicp_nn_hook hook,
                                void* user) {
  if (hook == nullptr) { dicp::set_error("dicp_icp_forward_hooked: null hook"); return DICP_ERR_ARG; }
  return dicp::icp_forward_impl<double>(source, target, T_init, w_init, m_valid, N, n, m, S, params, pc, T_out, deltas, weights, costs,
                                        converged, num_iters, match_ratio, saved_idx, saved_state, iters_run_host, ws, ws_bytes, stream, hook, user);
}
int dicp_solve_update_f32(const double* partials, int nblk, const double* C_r, int N, const dicp_params* params, float* delta,
                          double* C_r_next, uint8_t* below_tolerance, void* stream) {
  return dicp::solve_update_api<float>(partials, nblk, C_r, N, params, delta, C_r_next, below_tolerance, stream);
}
int dicp_solve_update_f64(const double* partials, int nblk, const double* C_r, int N, const dicp_params* params, double* delta,
                          double* C_r_next, uint8_t* below_tolerance, void* stream) {
  return dicp::solve_update_api<double>(partials, nblk, C_r, N, params, delta, C_r_next, below_tolerance, stream);
}
int dicp_gn_blocks_per_pair(int n) { return dicp::gn_blocks_per_pair(n); }
size_t dicp_gn_rows_bytes(int N, int m, int elem_size) { return dicp::gn_rows_bytes(N, m, elem_size); }
int dicp_gn_pack_rows_f32(const float* target, int N, int m, int S, int icp_type, void* rows, void* stream) {
  if (!target || !rows || N < 1 || m < 1 || (S != 3 && S != 6) || (icp_type == DICP_PT2PL && S != 6)) {
    dicp::set_error("dicp_gn_pack_rows: bad arguments");
    return DICP_ERR_ARG;
  }
  return dicp::gn_pack_rows<float>(target, N, m, S, icp_type == DICP_PT2PL, rows, (cudaStream_t)stream);
}
int dicp_gn_accumulate_f32(const float* source, const float* target, const void* rows, const float* w_init, const int32_t* idx,
                           const double* C_r, int N, int n, int m, int S, const dicp_params* params, float* weights_out,
                           double* partials_out, void* stream) {
  return dicp::gn_accumulate_api<float>(source, target, rows, w_init, idx, C_r, N, n, m, S, params, weights_out, partials_out, stream);
}
int dicp_gn_accumulate_f64(const double* source, const double* target, const double* w_init, const int32_t* idx, const double* C_r,
                           int N, int n, int m, int S, const dicp_params* params, double* weights_out, double* partials_out,
                           void* stream) {
  return dicp::gn_accumulate_api<double>(source, target, nullptr, w_init, idx, C_r, N, n, m, S, params, weights_out, partials_out, stream);
}
}
