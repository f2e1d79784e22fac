// kabsch_kernels.cu — the Kabsch / SVD point-to-point path (ICP.py:533-591).
//
// Stage kernels (dicp_kabsch_sums_*): centroids of the (moved) source points and of their nearest neighbours and the
// 3x3 cross-covariance as ONE fused pass over the correspondences (ICP.py:557-563):
//   sums[b] = ( sum x , sum q , sum q x^T ),   q_i = target[b, idx[b,i], :3],  x_i = C_a p_i + r_a
// from which  x_mean = Sx/n, q_mean = Sq/n, W = M/n - q_mean x_mean^T.  HBM bound: 12 (p) + 4 (idx) + 12 (gathered q) =
// 28 B per correspondence.  Their adjoint scatter-adds into the target.
//
// The whole loop (dicp_kabsch_icp_forward_* / _backward_*, second half of this file): search, sums, one thread per pair
// for the 3x3 SVD + determinant fix + composition (kabsch_step in dicp_math.cuh), device-side convergence test, and the
// hand-written reverse pass through the saved correspondences and per-iteration records.
#include "common.cuh"

namespace dicp {

constexpr int KB_THREADS = 128;
constexpr int KB_NSUM = 16;     // 3 + 3 + 9 (+1 pad)

template <typename T>
__global__ void __launch_bounds__(KB_THREADS)
kabsch_accum_kernel(const T* __restrict__ source, const T* __restrict__ target, const int32_t* __restrict__ idx, int n, int m, int S,
                    double* __restrict__ partials, const double* __restrict__ state, const int* __restrict__ stop_flag) {
  if (stop_flag != nullptr && *stop_flag != 0) return;
  __shared__ double smem[(KB_THREADS / 32) * KB_NSUM];
  constexpr int U = 4;
  const int b = blockIdx.y;
  // state (may be NULL): records whose accumulated (C_a, r_a) move the source points first, x = C_a p + r_a (ICP.py:581)
  T Ca[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, ra[3] = {0, 0, 0};
  if (state != nullptr) {
    const double* rec = state + (size_t)b * ST_STRIDE;
#pragma unroll
    for (int i = 0; i < 9; ++i) Ca[i] = (T)rec[KS_CA + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) ra[i] = (T)rec[KS_RA + i];
  }
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
      const T p0[3] = {src_b[3 * ic], src_b[3 * ic + 1], src_b[3 * ic + 2]};
      if (state != nullptr) {
        T a[3];
        rotate(Ca, p0, a);
        p[u][0] = a[0] + ra[0]; p[u][1] = a[1] + ra[1]; p[u][2] = a[2] + ra[2];
      } else {
        p[u][0] = p0[0]; p[u][1] = p0[1]; p[u][2] = p0[2];
      }
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
  kabsch_accum_kernel<T><<<dim3(nblk, N), KB_THREADS, 0, st>>>(source, target, idx, n, m, S, partials, nullptr, nullptr);
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

// =========================================================================================
// The whole Kabsch / SVD point-to-point loop (ICP.py:533-591) on the device.
// Per iteration: nearest-neighbour search of the moved points (the search kernels apply the accumulated
// (C_a, r_a) of the iteration's record), fused centroid / cross-covariance sums, one thread per pair for the
// 3x3 SVD + determinant fix + composition (kabsch_step, dicp_math.cuh), and -- when a tolerance is set -- the
// residual pass of the convergence test (ICP.py:585).  The reverse pass is hand-written from the saved
// correspondences and per-iteration records (U, s, V, centroids, C, r).
// Reference quirks kept: T_init is only composed into the returned transform, the source is not moved by it
// before the first search (ICP.py:545-547, so T = [C_a r_a] T_init); convergence compares the moved points
// with the neighbours found before the move.
// =========================================================================================
struct KbCtrl { int stop, iters_run; };

template <typename T>
__global__ void kabsch_init_kernel(int N, double* __restrict__ state0, KbCtrl* __restrict__ ctrl) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b == 0) { ctrl->stop = 0; ctrl->iters_run = 0; }
  if (b >= N) return;
  double* rec = state0 + (size_t)b * ST_STRIDE;
  for (int i = 0; i < ST_STRIDE; ++i) rec[i] = 0.0;
  rec[KS_CA] = rec[KS_CA + 4] = rec[KS_CA + 8] = 1.0;
}

// sums -> SVD step -> next record's accumulated transform
template <typename T>
__global__ void kabsch_update_kernel(const double* __restrict__ partials, int nblk, double* __restrict__ rec_k, double* __restrict__ rec_next,
                                     int N, int n, int literal, int k, KbCtrl* __restrict__ ctrl, int count_iteration) {
  if (ctrl->stop != 0) return;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= N) return;
  double sums[KB_NSUM];
  for (int v = 0; v < 15; ++v) {
    double s = 0.0;
    for (int j = 0; j < nblk; ++j) s += partials[((size_t)b * nblk + j) * KB_NSUM + v];      // fixed order: deterministic
    sums[v] = s;
  }
  double* rec = rec_k + (size_t)b * ST_STRIDE;
  double* nxt = rec_next + (size_t)b * ST_STRIDE;
  kabsch_step(sums, (double)n, literal != 0, rec);
  const double* C = rec + KS_C; const double* r = rec + KS_R; const double* Ca = rec + KS_CA; const double* ra = rec + KS_RA;
  for (int i = 0; i < ST_STRIDE; ++i) nxt[i] = 0.0;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j)                                                                  // ps <- C ps + r (ICP.py:581), accumulated
      nxt[KS_CA + i * 3 + j] = (double)(T)(C[i * 3] * Ca[j] + C[i * 3 + 1] * Ca[3 + j] + C[i * 3 + 2] * Ca[6 + j]);
    nxt[KS_RA + i] = (double)(T)(C[i * 3] * ra[0] + C[i * 3 + 1] * ra[1] + C[i * 3 + 2] * ra[2] + r[i]);
  }
  if (count_iteration && b == 0) ctrl->iters_run = k + 1;      // no convergence test in this call: every launched iteration counts
}

// sum |x_i - q_i|^2 with the points moved by the NEXT record's transform and the neighbours of this iteration (ICP.py:585)
template <typename T>
__global__ void __launch_bounds__(KB_THREADS)
kabsch_residual_kernel(const T* __restrict__ source, const T* __restrict__ target, const int32_t* __restrict__ idx, int n, int m, int S,
                       const double* __restrict__ rec_next, double* __restrict__ partials, const KbCtrl* __restrict__ ctrl) {
  if (ctrl->stop != 0) return;
  __shared__ double smem[KB_THREADS / 32];
  const int b = blockIdx.y;
  const double* rec = rec_next + (size_t)b * ST_STRIDE;
  T Ca[9], ra[3];
#pragma unroll
  for (int i = 0; i < 9; ++i) Ca[i] = (T)rec[KS_CA + i];
#pragma unroll
  for (int i = 0; i < 3; ++i) ra[i] = (T)rec[KS_RA + i];
  T acc = T(0);
  for (int i = blockIdx.x * KB_THREADS + threadIdx.x; i < n; i += gridDim.x * KB_THREADS) {
    const T* p = source + ((size_t)b * n + i) * 3;
    const T pv[3] = {p[0], p[1], p[2]};
    T a[3];
    rotate(Ca, pv, a);
    const T* row = target + ((size_t)b * m + idx[(size_t)b * n + i]) * S;
    const T e0 = a[0] + ra[0] - row[0], e1 = a[1] + ra[1] - row[1], e2 = a[2] + ra[2] - row[2];
    acc += e0 * e0 + e1 * e1 + e2 * e2;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = (double)acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < KB_THREADS / 32; ++w) s += smem[w];
    partials[(size_t)b * gridDim.x + blockIdx.x] = s;
  }
}

// one CTA: per-pair residuals, "every pair below the tolerance" -> stop (batched form of the break at ICP.py:585-586)
__global__ void __launch_bounds__(256)
kabsch_check_kernel(const double* __restrict__ partials, int nblk, double* __restrict__ rec_k, int N, double tolerance, int k,
                    KbCtrl* __restrict__ ctrl) {
  if (ctrl->stop != 0) return;
  __shared__ int s_open;
  if (threadIdx.x == 0) s_open = 0;
  __syncthreads();
  for (int b = threadIdx.x; b < N; b += blockDim.x) {
    double s = 0.0;
    for (int j = 0; j < nblk; ++j) s += partials[(size_t)b * nblk + j];
    rec_k[(size_t)b * ST_STRIDE + KS_RES] = s;
    if (!(s < tolerance)) atomicExch(&s_open, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    ctrl->iters_run = k + 1;
    if (s_open == 0) ctrl->stop = 1;
  }
}

// pc = C_a p + r_a with the final accumulated transform, T = [C_a r_a; 0 1] T_init (ICP.py:578, 581)
template <typename T>
__global__ void __launch_bounds__(256)
kabsch_finalize_kernel(const T* __restrict__ source, const T* __restrict__ T_init, int n, const double* __restrict__ state_all, int N,
                       const KbCtrl* __restrict__ ctrl, T* __restrict__ pc, T* __restrict__ T_out) {
  const int b = blockIdx.y;
  const double* rec = state_all + ((size_t)ctrl->iters_run * N + b) * ST_STRIDE;
  T Ca[9], ra[3];
#pragma unroll
  for (int i = 0; i < 9; ++i) Ca[i] = (T)rec[KS_CA + i];
#pragma unroll
  for (int i = 0; i < 3; ++i) ra[i] = (T)rec[KS_RA + i];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const T* p = source + ((size_t)b * n + i) * 3;
    const T pv[3] = {p[0], p[1], p[2]};
    T a[3];
    rotate(Ca, pv, a);
    T* o = pc + ((size_t)b * n + i) * 3;
    o[0] = a[0] + ra[0]; o[1] = a[1] + ra[1]; o[2] = a[2] + ra[2];
  }
  if (blockIdx.x == 0 && threadIdx.x < 16) {
    const int rr = threadIdx.x >> 2, c = threadIdx.x & 3;
    const T* Ti = T_init + (size_t)b * 16;
    double v;
    if (rr < 3) v = rec[KS_CA + rr * 3] * (double)Ti[c] + rec[KS_CA + rr * 3 + 1] * (double)Ti[4 + c] + rec[KS_CA + rr * 3 + 2] * (double)Ti[8 + c] +
                    rec[KS_RA + rr] * (double)Ti[12 + c];
    else v = (double)Ti[12 + c];
    T_out[(size_t)b * 16 + threadIdx.x] = (T)v;
  }
}

struct KabschWorkspace {
  void* packed; void* nn_ws; size_t nn_ws_bytes;
  double* partials; double* res_partials; KbCtrl* ctrl; double* bw_state; double* bw_partials;
  size_t total;
};
constexpr int KBW_STRIDE = 32;      // backward carry per pair: Cbar_a 9, rbar_a 3, g 15 (adjoints of the sums)
constexpr int KBW_G = 12;
static KabschWorkspace kabsch_plan(void* ws, size_t ws_bytes, int N, int n, int m, int elem_size) {
  Arena ar(ws, ws_bytes);
  KabschWorkspace w;
  const int nblk = kb_blocks(n);
  w.packed = ar.take<char>(nn_packed_bytes(N, m, elem_size));
  w.nn_ws_bytes = nn_search_ws_bytes(N, n, m, elem_size);
  w.nn_ws = ar.take<char>(w.nn_ws_bytes > 0 ? w.nn_ws_bytes : 1);
  w.partials = ar.take<double>((size_t)N * nblk * KB_NSUM);
  w.res_partials = ar.take<double>((size_t)N * nblk);
  w.ctrl = ar.take<KbCtrl>(1);
  w.bw_state = ar.take<double>((size_t)N * KBW_STRIDE);
  w.bw_partials = ar.take<double>((size_t)N * 2 * nblk * BW_NRED);
  w.total = ar.off;
  return w;
}

template <typename T>
static int kabsch_icp_forward_impl(const T* source, const T* target, const T* T_init, int N, int n, int m, int S, int max_iterations,
                                   double tolerance, int literal, int check_every, T* pc, T* T_out, int32_t* saved_idx,
                                   double* saved_state, int32_t* iters_run_host, void* ws, size_t ws_bytes, void* stream) {
  DICP_CHECK_ARG(source && target && T_init && pc && T_out && saved_idx && saved_state && iters_run_host, "null pointer");
  DICP_CHECK_ARG(N >= 1 && N <= 65535 && n >= 1 && m >= 1 && (S == 3 || S == 6) && max_iterations >= 1, "bad sizes");
  if (ws == nullptr || ws_bytes < kabsch_plan(nullptr, 0, N, n, m, sizeof(T)).total) {
    set_error("dicp_kabsch_icp_forward: workspace too small");
    return DICP_ERR_WORKSPACE;
  }
  typedef typename Vec4<T>::type V4;
  cudaStream_t st = (cudaStream_t)stream;
  const KabschWorkspace w = kabsch_plan(ws, ws_bytes, N, n, m, sizeof(T));
  const int nblk = kb_blocks(n), K = max_iterations;
  const size_t rec_stride = (size_t)N * ST_STRIDE;
  const bool test = tolerance > 0.0;            // sum of squares < tolerance can never hold otherwise (ICP.py:585)
  if (check_every < 1) check_every = 1;
  kabsch_init_kernel<T><<<(N + 127) / 128, 128, 0, st>>>(N, saved_state, w.ctrl);
  DICP_LAUNCH_CHECK();
  int rc = nn_pack_targets<T>(target, N, m, S, nullptr, (V4*)w.packed, st);
  if (rc) return rc;
  KbCtrl h = {0, 0};
  for (int k = 0; k < K; ++k) {
    double* rec_k = saved_state + (size_t)k * rec_stride;
    int32_t* idx_k = saved_idx + (size_t)k * N * n;
    rc = nn_search<T>(source, rec_k, (const V4*)w.packed, N, n, m, idx_k, nullptr, &w.ctrl->stop, w.nn_ws, w.nn_ws_bytes, st);   // ICP.py:553-555
    if (rc) return rc;
    kabsch_accum_kernel<T><<<dim3(nblk, N), KB_THREADS, 0, st>>>(source, target, idx_k, n, m, S, w.partials, rec_k, &w.ctrl->stop);
    DICP_LAUNCH_CHECK();
    kabsch_update_kernel<T><<<(N + 63) / 64, 64, 0, st>>>(w.partials, nblk, rec_k, rec_k + rec_stride, N, n, literal, k, w.ctrl, test ? 0 : 1);
    DICP_LAUNCH_CHECK();
    if (test) {
      kabsch_residual_kernel<T><<<dim3(nblk, N), KB_THREADS, 0, st>>>(source, target, idx_k, n, m, S, rec_k + rec_stride, w.res_partials, w.ctrl);
      DICP_LAUNCH_CHECK();
      kabsch_check_kernel<<<1, 256, 0, st>>>(w.res_partials, nblk, rec_k, N, tolerance, k, w.ctrl);
      DICP_LAUNCH_CHECK();
      if ((k + 1) % check_every == 0 && k + 1 < K) {
        DICP_CUDA(cudaMemcpyAsync(&h, w.ctrl, sizeof(KbCtrl), cudaMemcpyDeviceToHost, st));
        DICP_CUDA(cudaStreamSynchronize(st));
        if (h.stop) break;
      }
    }
  }
  kabsch_finalize_kernel<T><<<dim3(kb_blocks(n), N), 256, 0, st>>>(source, T_init, n, saved_state, N, w.ctrl, pc, T_out);
  DICP_LAUNCH_CHECK();
  DICP_CUDA(cudaMemcpyAsync(&h, w.ctrl, sizeof(KbCtrl), cudaMemcpyDeviceToHost, st));
  DICP_CUDA(cudaStreamSynchronize(st));
  *iters_run_host = h.iters_run;
  return 0;
}

// ---- reverse pass -----------------------------------------------------------------------------
// seed: pc = C_a p + r_a  ->  grad_source = C_a^T pcbar, partial sums of sum(pcbar p^T), sum(pcbar)
template <typename T>
__global__ void __launch_bounds__(KB_THREADS)
kabsch_bwd_seed_kernel(const T* __restrict__ source, const T* __restrict__ grad_pc, const double* __restrict__ rec_final, int n,
                       T* __restrict__ grad_source, double* __restrict__ partials) {
  __shared__ double smem[(KB_THREADS / 32) * BW_NRED];
  const int b = blockIdx.y;
  const double* rec = rec_final + (size_t)b * ST_STRIDE;
  T Ca[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) Ca[i] = (T)rec[KS_CA + i];
  T acc[BW_NRED];
#pragma unroll
  for (int i = 0; i < BW_NRED; ++i) acc[i] = T(0);
  for (int i = blockIdx.x * KB_THREADS + threadIdx.x; i < n; i += gridDim.x * KB_THREADS) {
    T g[3] = {T(0), T(0), T(0)};
    if (grad_pc != nullptr) {
      const T* gp = grad_pc + ((size_t)b * n + i) * 3;
      g[0] = gp[0]; g[1] = gp[1]; g[2] = gp[2];
    }
    const T* p = source + ((size_t)b * n + i) * 3;
    T* gs = grad_source + ((size_t)b * n + i) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) gs[c] = Ca[c] * g[0] + Ca[3 + c] * g[1] + Ca[6 + c] * g[2];
#pragma unroll
    for (int rr = 0; rr < 3; ++rr) {
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[rr * 3 + c] += g[rr] * p[c];
      acc[9 + rr] += g[rr];
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int v = 0; v < BW_NRED; ++v) {
    T x = acc[v];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (lane == 0) smem[warp * BW_NRED + v] = (double)x;
  }
  __syncthreads();
  if (threadIdx.x < BW_NRED) {
    double s = 0.0;
    for (int w = 0; w < KB_THREADS / 32; ++w) s += smem[w * BW_NRED + threadIdx.x];
    partials[((size_t)b * gridDim.x + blockIdx.x) * BW_NRED + threadIdx.x] = s;
  }
}

// per pair: mode 0 starts the carry from grad_T (T = A T_init => Abar = Tbar T_init^T, grad_T_init = A^T Tbar), every mode adds
// the partial sums of the previous point pass, then the adjoint of iteration k's composition and SVD step.
template <typename T>
__global__ void kabsch_bwd_pair_kernel(const double* __restrict__ partials, int nblk, const double* __restrict__ rec_k,
                                       const double* __restrict__ rec_final, int n, int literal, int mode, const T* __restrict__ grad_T,
                                       const T* __restrict__ T_init, double* __restrict__ bw_state, T* __restrict__ grad_T_init, int N) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= N) return;
  double* bw = bw_state + (size_t)b * KBW_STRIDE;
  double carry[12];
  if (mode == 0) {
    const double* fin = rec_final + (size_t)b * ST_STRIDE;
    double Tb[16], Ti[16];
    for (int i = 0; i < 16; ++i) { Tb[i] = grad_T != nullptr ? (double)grad_T[(size_t)b * 16 + i] : 0.0; Ti[i] = (double)T_init[(size_t)b * 16 + i]; }
    for (int rr = 0; rr < 3; ++rr) {
      for (int c = 0; c < 3; ++c) carry[rr * 3 + c] = Tb[rr * 4] * Ti[c * 4] + Tb[rr * 4 + 1] * Ti[c * 4 + 1] + Tb[rr * 4 + 2] * Ti[c * 4 + 2] + Tb[rr * 4 + 3] * Ti[c * 4 + 3];
      carry[9 + rr] = Tb[rr * 4] * Ti[12] + Tb[rr * 4 + 1] * Ti[13] + Tb[rr * 4 + 2] * Ti[14] + Tb[rr * 4 + 3] * Ti[15];
    }
    // grad_T_init = A^T Tbar, A = [C_a r_a; 0 1]
    T* g0 = grad_T_init + (size_t)b * 16;
    for (int c = 0; c < 4; ++c) {
      for (int rr = 0; rr < 3; ++rr) g0[rr * 4 + c] = (T)(fin[KS_CA + rr] * Tb[c] + fin[KS_CA + 3 + rr] * Tb[4 + c] + fin[KS_CA + 6 + rr] * Tb[8 + c]);
      g0[12 + c] = (T)(fin[KS_RA] * Tb[c] + fin[KS_RA + 1] * Tb[4 + c] + fin[KS_RA + 2] * Tb[8 + c] + Tb[12 + c]);
    }
  } else {
    for (int i = 0; i < 12; ++i) carry[i] = bw[i];
  }
  for (int j = 0; j < nblk; ++j)
    for (int i = 0; i < 12; ++i) carry[i] += partials[((size_t)b * nblk + j) * BW_NRED + i];
  const double* rec = rec_k + (size_t)b * ST_STRIDE;
  const double* C = rec + KS_C; const double* Ca = rec + KS_CA; const double* ra = rec + KS_RA;
  // C_a' = C C_a, r_a' = C r_a + r
  double Cbar[9], rbar[3], nCa[9], nra[3];
  for (int i = 0; i < 3; ++i) {
    rbar[i] = carry[9 + i];
    for (int j = 0; j < 3; ++j) {
      Cbar[i * 3 + j] = carry[i * 3] * Ca[j * 3] + carry[i * 3 + 1] * Ca[j * 3 + 1] + carry[i * 3 + 2] * Ca[j * 3 + 2] + carry[9 + i] * ra[j];
      nCa[i * 3 + j] = C[i] * carry[j] + C[3 + i] * carry[3 + j] + C[6 + i] * carry[6 + j];
    }
    nra[i] = C[i] * carry[9] + C[3 + i] * carry[10] + C[6 + i] * carry[11];
  }
  double g[15];
  kabsch_step_backward(rec, (double)n, literal != 0, Cbar, rbar, g);
  for (int i = 0; i < 9; ++i) bw[i] = nCa[i];
  for (int i = 0; i < 3; ++i) bw[9 + i] = nra[i];
  for (int i = 0; i < 15; ++i) bw[KBW_G + i] = g[i];
}

// per correspondence: x = C_a p + r_a, q = target[idx]; xbar = gSx + gM^T q, qbar = gSq + gM x (scatter-add),
// grad_source += C_a^T xbar, partial sums of sum(xbar p^T), sum(xbar)
template <typename T>
__global__ void __launch_bounds__(KB_THREADS)
kabsch_bwd_points_kernel(const T* __restrict__ source, const T* __restrict__ target, const int32_t* __restrict__ idx, int n, int m, int S,
                         const double* __restrict__ rec_k, const double* __restrict__ bw_state, T* __restrict__ grad_source,
                         T* __restrict__ grad_target, double* __restrict__ partials) {
  __shared__ double smem[(KB_THREADS / 32) * BW_NRED];
  const int b = blockIdx.y;
  const double* rec = rec_k + (size_t)b * ST_STRIDE;
  const double* bw = bw_state + (size_t)b * KBW_STRIDE;
  T Ca[9], ra[3], g[15];
#pragma unroll
  for (int i = 0; i < 9; ++i) Ca[i] = (T)rec[KS_CA + i];
#pragma unroll
  for (int i = 0; i < 3; ++i) ra[i] = (T)rec[KS_RA + i];
#pragma unroll
  for (int i = 0; i < 15; ++i) g[i] = (T)bw[KBW_G + i];
  T acc[BW_NRED];
#pragma unroll
  for (int i = 0; i < BW_NRED; ++i) acc[i] = T(0);
  for (int i = blockIdx.x * KB_THREADS + threadIdx.x; i < n; i += gridDim.x * KB_THREADS) {
    const T* pp = source + ((size_t)b * n + i) * 3;
    const T p[3] = {pp[0], pp[1], pp[2]};
    T a[3];
    rotate(Ca, p, a);
    const T x[3] = {a[0] + ra[0], a[1] + ra[1], a[2] + ra[2]};
    const int j = idx[(size_t)b * n + i];
    const T* row = target + ((size_t)b * m + j) * S;
    const T q[3] = {row[0], row[1], row[2]};
    T xb[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) xb[c] = g[c] + g[6 + c] * q[0] + g[9 + c] * q[1] + g[12 + c] * q[2];
    T* gt = grad_target + ((size_t)b * m + j) * S;
#pragma unroll
    for (int rr = 0; rr < 3; ++rr) atomicAdd(gt + rr, g[3 + rr] + g[6 + rr * 3] * x[0] + g[7 + rr * 3] * x[1] + g[8 + rr * 3] * x[2]);
    T* gs = grad_source + ((size_t)b * n + i) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) gs[c] += Ca[c] * xb[0] + Ca[3 + c] * xb[1] + Ca[6 + c] * xb[2];
#pragma unroll
    for (int rr = 0; rr < 3; ++rr) {
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[rr * 3 + c] += xb[rr] * p[c];
      acc[9 + rr] += xb[rr];
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int v = 0; v < BW_NRED; ++v) {
    T x = acc[v];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (lane == 0) smem[warp * BW_NRED + v] = (double)x;
  }
  __syncthreads();
  if (threadIdx.x < BW_NRED) {
    double s = 0.0;
    for (int w = 0; w < KB_THREADS / 32; ++w) s += smem[w * BW_NRED + threadIdx.x];
    partials[((size_t)b * gridDim.x + blockIdx.x) * BW_NRED + threadIdx.x] = s;
  }
}

template <typename T>
static int kabsch_icp_backward_impl(const T* source, const T* target, const T* T_init, int N, int n, int m, int S, int literal,
                                    int iters_run, const int32_t* saved_idx, const double* saved_state, const T* grad_T,
                                    const T* grad_pc, T* grad_source, T* grad_target, T* grad_T_init, void* ws, size_t ws_bytes,
                                    void* stream) {
  DICP_CHECK_ARG(source && target && T_init && saved_idx && saved_state && grad_source && grad_target && grad_T_init, "null pointer");
  DICP_CHECK_ARG(N >= 1 && N <= 65535 && n >= 1 && m >= 1 && (S == 3 || S == 6) && iters_run >= 1, "bad sizes");
  if (ws == nullptr || ws_bytes < kabsch_plan(nullptr, 0, N, n, m, sizeof(T)).total) {
    set_error("dicp_kabsch_icp_backward: workspace too small");
    return DICP_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const KabschWorkspace w = kabsch_plan(ws, ws_bytes, N, n, m, sizeof(T));
  const int nblk = 2 * kb_blocks(n);
  const size_t rec_stride = (size_t)N * ST_STRIDE;
  const double* rec_final = saved_state + (size_t)iters_run * rec_stride;
  DICP_CUDA(cudaMemsetAsync(grad_target, 0, (size_t)N * m * S * sizeof(T), st));
  kabsch_bwd_seed_kernel<T><<<dim3(nblk, N), KB_THREADS, 0, st>>>(source, grad_pc, rec_final, n, grad_source, w.bw_partials);
  DICP_LAUNCH_CHECK();
  for (int k = iters_run - 1; k >= 0; --k) {
    const double* rec_k = saved_state + (size_t)k * rec_stride;
    kabsch_bwd_pair_kernel<T><<<(N + 63) / 64, 64, 0, st>>>(w.bw_partials, nblk, rec_k, rec_final, n, literal, k == iters_run - 1 ? 0 : 1, grad_T,
                                                            T_init, w.bw_state, grad_T_init, N);
    DICP_LAUNCH_CHECK();
    kabsch_bwd_points_kernel<T><<<dim3(nblk, N), KB_THREADS, 0, st>>>(source, target, saved_idx + (size_t)k * N * n, n, m, S, rec_k, w.bw_state,
                                                                      grad_source, grad_target, w.bw_partials);
    DICP_LAUNCH_CHECK();
  }
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
size_t dicp_kabsch_icp_workspace_bytes(int N, int n, int m, int elem_size) { return dicp::kabsch_plan(nullptr, 0, N, n, m, elem_size).total; }
int dicp_kabsch_icp_forward_f32(const float* source, const float* target, const float* T_init, int N, int n, int m, int S, int max_iterations,
                                double tolerance, int literal, int check_every, float* pc, float* T_out, int32_t* saved_idx,
                                double* saved_state, int32_t* iters_run_host, void* ws, size_t ws_bytes, void* stream) {
  return dicp::kabsch_icp_forward_impl<float>(source, target, T_init, N, n, m, S, max_iterations, tolerance, literal, check_every, pc, T_out,
                                              saved_idx, saved_state, iters_run_host, ws, ws_bytes, stream);
}
int dicp_kabsch_icp_forward_f64(const double* source, const double* target, const double* T_init, int N, int n, int m, int S, int max_iterations,
                                double tolerance, int literal, int check_every, double* pc, double* T_out, int32_t* saved_idx,
                                double* saved_state, int32_t* iters_run_host, void* ws, size_t ws_bytes, void* stream) {
  return dicp::kabsch_icp_forward_impl<double>(source, target, T_init, N, n, m, S, max_iterations, tolerance, literal, check_every, pc, T_out,
                                               saved_idx, saved_state, iters_run_host, ws, ws_bytes, stream);
}
int dicp_kabsch_icp_backward_f32(const float* source, const float* target, const float* T_init, int N, int n, int m, int S, int literal,
                                 int iters_run, const int32_t* saved_idx, const double* saved_state, const float* grad_T, const float* grad_pc,
                                 float* grad_source, float* grad_target, float* grad_T_init, void* ws, size_t ws_bytes, void* stream) {
  return dicp::kabsch_icp_backward_impl<float>(source, target, T_init, N, n, m, S, literal, iters_run, saved_idx, saved_state, grad_T, grad_pc,
                                               grad_source, grad_target, grad_T_init, ws, ws_bytes, stream);
}
int dicp_kabsch_icp_backward_f64(const double* source, const double* target, const double* T_init, int N, int n, int m, int S, int literal,
                                 int iters_run, const int32_t* saved_idx, const double* saved_state, const double* grad_T, const double* grad_pc,
                                 double* grad_source, double* grad_target, double* grad_T_init, void* ws, size_t ws_bytes, void* stream) {
  return dicp::kabsch_icp_backward_impl<double>(source, target, T_init, N, n, m, S, literal, iters_run, saved_idx, saved_state, grad_T, grad_pc,
                                                grad_source, grad_target, grad_T_init, ws, ws_bytes, stream);
}
}
