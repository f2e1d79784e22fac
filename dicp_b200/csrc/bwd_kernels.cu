// bwd_kernels.cu — hand-written reverse pass of the dICP loop (replaces torch autograd
// through ICP.py:132-280; derivation and executable spec: oracle/adjoint.py).
//
// The forward saved, per iteration k: the correspondences idx_k [N,n] and a per-pair
// record (C_k, r_k, active_k, delta_k, A_k).  Everything per-point is recomputed.
// Reverse order, per iteration two launches:
//   bwd_pair_kernel    per pair: fold the point pass's CTA partial sums into (Cbar, rbar),
//                      push them through the update step (C+ = R^T C, r+ = r - rho),
//                      lambda = A^-1 [phibar; -rbar]   (6x6 pivoted solve in double)
//   bwd_points_kernel  per correspondence: adjoints of residual / weights / Jacobian;
//                      += into grad_source and grad_w (private rows, plain RMW),
//                      atomicAdd into grad_target[idx] (the scatter-add adjoint of the
//                      gather, nn.py:37-38), CTA partial sums of sum(abar p^T), sum(ebar).
// HBM-bound: 124 B per correspondence (pt2pl).
#include "common.cuh"

namespace dicp {

constexpr int BW_THREADS = 128;

template <typename T, int NV>
__device__ __forceinline__ void bw_block_reduce_store(const T* vals, double* out, double* smem) {
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

// seeds from pc = C_K p + r_K (ICP.py:274): grad_source = C_K^T pcbar, partial sums of
// sum(pcbar p^T) and sum(pcbar)
template <typename T>
__global__ void __launch_bounds__(BW_THREADS)
bwd_seed_kernel(const T* __restrict__ source, const T* __restrict__ grad_pc, const double* __restrict__ state_final, int n,
                T* __restrict__ grad_source, double* __restrict__ partials) {
  __shared__ double red_smem[(BW_THREADS / 32) * BW_NRED];
  const int b = blockIdx.y;
  const double* rec = state_final + (size_t)b * ST_STRIDE;
  T C[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) C[i] = (T)rec[ST_C + i];
  T acc[BW_NRED];
#pragma unroll
  for (int i = 0; i < BW_NRED; ++i) acc[i] = T(0);
  const T* src_b = source + (size_t)b * n * 3;
  T* gs_b = grad_source + (size_t)b * n * 3;
  for (int i = blockIdx.x * BW_THREADS + threadIdx.x; i < n; i += gridDim.x * BW_THREADS) {
    T g[3] = {T(0), T(0), T(0)};
    if (grad_pc != nullptr) {
      const T* gp = grad_pc + ((size_t)b * n + i) * 3;
      g[0] = gp[0]; g[1] = gp[1]; g[2] = gp[2];
    }
    const T p[3] = {src_b[3 * i], src_b[3 * i + 1], src_b[3 * i + 2]};
#pragma unroll
    for (int c = 0; c < 3; ++c) gs_b[3 * i + c] = C[c] * g[0] + C[3 + c] * g[1] + C[6 + c] * g[2];
#pragma unroll
    for (int rr = 0; rr < 3; ++rr) {
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[rr * 3 + c] += g[rr] * p[c];
      acc[9 + rr] += g[rr];
    }
  }
  bw_block_reduce_store<T, BW_NRED>(acc, partials + ((size_t)b * gridDim.x + blockIdx.x) * BW_NRED, red_smem);
}

// mode 0: first step (k = K-1): carry = grad_T + partial sums, then update-step adjoint
// mode 1: later steps:          carry += partial sums,          then update-step adjoint
// mode 2: after k = 0:          carry += partial sums, write grad_T_init
template <typename T>
__global__ void bwd_pair_kernel(const double* __restrict__ partials, int nblk, const double* __restrict__ rec_k, int dim, int mode,
                                const T* __restrict__ grad_T, double* __restrict__ bw_state, T* __restrict__ grad_T_init, int N) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= N) return;
  double* bw = bw_state + (size_t)b * BW_STRIDE;
  double carry[12];
  if (mode == 0) {
    for (int rr = 0; rr < 3; ++rr) {
      for (int c = 0; c < 3; ++c) carry[rr * 3 + c] = grad_T != nullptr ? (double)grad_T[(size_t)b * 16 + rr * 4 + c] : 0.0;
      carry[9 + rr] = grad_T != nullptr ? (double)grad_T[(size_t)b * 16 + rr * 4 + 3] : 0.0;
    }
  } else {
    for (int i = 0; i < 12; ++i) carry[i] = bw[i];
  }
  for (int j = 0; j < nblk; ++j)
    for (int i = 0; i < 12; ++i) carry[i] += partials[((size_t)b * nblk + j) * BW_NRED + i];
  if (mode == 2) {
    T* g = grad_T_init + (size_t)b * 16;
    for (int rr = 0; rr < 3; ++rr) {
      for (int c = 0; c < 3; ++c) g[rr * 4 + c] = (T)carry[rr * 3 + c];
      g[rr * 4 + 3] = (T)carry[9 + rr];
    }
    g[12] = g[13] = g[14] = g[15] = T(0);
    return;
  }
  const double* rec = rec_k + (size_t)b * ST_STRIDE;
  double lam[6] = {0, 0, 0, 0, 0, 0};
  const double active = rec[ST_ACTIVE];
  if (active != 0.0) pair_backward(rec, dim, carry + BW_CBAR, carry + BW_RBAR, lam);
  // frozen pair: w = 0 => A = 1e-12 I, delta = 0, R = I; nothing flows into the points and
  // (Cbar, rbar) pass through unchanged (ICP.py:256-257)
  for (int i = 0; i < 12; ++i) bw[i] = carry[i];
  for (int i = 0; i < 6; ++i) { bw[BW_LAM + i] = lam[i]; bw[BW_DEL + i] = rec[ST_DELTA + i]; }
  bw[BW_ACTIVE] = active;
}

template <typename T, bool PT2PL, bool FASTW>
__global__ void __launch_bounds__(BW_THREADS)
bwd_points_kernel(const T* __restrict__ source, const T* __restrict__ target, const T* __restrict__ w_init,
                  const int32_t* __restrict__ idx, const double* __restrict__ rec_k, const double* __restrict__ bw_state,
                  int n, int m, int S, WeightParams P, T* __restrict__ grad_source, T* __restrict__ grad_target,
                  T* __restrict__ grad_w, double* __restrict__ partials) {
  __shared__ double red_smem[(BW_THREADS / 32) * BW_NRED];
  const int b = blockIdx.y;
  const double* rec = rec_k + (size_t)b * ST_STRIDE;
  const double* bw = bw_state + (size_t)b * BW_STRIDE;
  T acc[BW_NRED];
#pragma unroll
  for (int i = 0; i < BW_NRED; ++i) acc[i] = T(0);
  if (bw[BW_ACTIVE] != 0.0) {
    T C[9], r[3], lam[6], del[6];
#pragma unroll
    for (int i = 0; i < 9; ++i) C[i] = (T)rec[ST_C + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) r[i] = (T)rec[ST_R + i];
#pragma unroll
    for (int i = 0; i < 6; ++i) { lam[i] = (T)bw[BW_LAM + i]; del[i] = (T)bw[BW_DEL + i]; }
    const T* src_b = source + (size_t)b * n * 3;
    const T* tgt_b = target + (size_t)b * m * S;
    T* gt_b = grad_target + (size_t)b * m * S;
    T* gs_b = grad_source + (size_t)b * n * 3;
    const int32_t* idx_b = idx + (size_t)b * n;
    // U correspondences per thread and pass, loads hoisted ahead of the arithmetic (see gn_accum_kernel)
    constexpr int U = sizeof(T) == 4 ? 4 : 2;
    for (int base = blockIdx.x * (BW_THREADS * U); base < n; base += gridDim.x * (BW_THREADS * U)) {
      int ii[U], jj[U];
      T p[U][3], q[U][3], nr[U][3], w0[U], gs[U][3], gw[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = base + u * BW_THREADS + threadIdx.x;
        ii[u] = i < n ? i : -1;
        const int ic = i < n ? i : n - 1;
        jj[u] = idx_b[ic];
        p[u][0] = src_b[3 * ic]; p[u][1] = src_b[3 * ic + 1]; p[u][2] = src_b[3 * ic + 2];
        w0[u] = w_init != nullptr ? w_init[(size_t)b * n + ic] : T(1);
        // the running gradients this pass adds to: read with the other loads, not after the arithmetic
        gs[u][0] = gs_b[3 * ic]; gs[u][1] = gs_b[3 * ic + 1]; gs[u][2] = gs_b[3 * ic + 2];
        gw[u] = grad_w != nullptr ? grad_w[(size_t)b * n + ic] : T(0);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        load_target_row<T, PT2PL>(tgt_b + (size_t)jj[u] * S, S, q[u], nr[u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (ii[u] < 0) continue;
        const int i = ii[u];
        PointEval<T> s;
        eval_point<T, PT2PL, FASTW>(C, r, p[u], q[u], nr[u], w0[u], P, s);
        T abar[3], ebar[3], nbar[3], w0bar;
        point_backward<T, PT2PL>(s, w0[u], lam, del, abar, ebar, nbar, w0bar);
        if (grad_w != nullptr) grad_w[(size_t)b * n + i] = gw[u] + w0bar;
#pragma unroll
        for (int c = 0; c < 3; ++c) gs_b[3 * i + c] = gs[u][c] + (C[c] * abar[0] + C[3 + c] * abar[1] + C[6 + c] * abar[2]);   // C^T abar
        T* grow = gt_b + (size_t)jj[u] * S;
        if (PT2PL) {
          const T g6[6] = {-ebar[0], -ebar[1], -ebar[2], nbar[0], nbar[1], nbar[2]};
          atomic_add_row6(grow, g6);
        } else {
#pragma unroll
          for (int c = 0; c < 3; ++c) atomicAdd(grow + c, -ebar[c]);
        }
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) {
#pragma unroll
          for (int c = 0; c < 3; ++c) acc[rr * 3 + c] += abar[rr] * p[u][c];
          acc[9 + rr] += ebar[rr];
        }
      }
    }
  }
  bw_block_reduce_store<T, BW_NRED>(acc, partials + ((size_t)b * gridDim.x + blockIdx.x) * BW_NRED, red_smem);
}

template <typename T>
static int icp_backward_impl(const T* source, const T* target, const T* w_init, int N, int n, int m, int S, const dicp_params* p,
                             int iters_run, const int32_t* saved_idx, const double* saved_state, const T* grad_T, const T* grad_pc,
                             T* grad_source, T* grad_target, T* grad_w, T* grad_T_init, void* ws, size_t ws_bytes, void* stream) {
  DICP_CHECK_ARG(source && target && p && saved_idx && saved_state && grad_source && grad_target && grad_T_init, "null pointer");
  DICP_CHECK_ARG(N >= 1 && N <= 65535 && n >= 1 && m >= 1 && (S == 3 || S == 6), "bad sizes");
  DICP_CHECK_ARG(p->icp_type != DICP_PT2PL || S == 6, "pt2pl needs a 6-column target");
  DICP_CHECK_ARG(iters_run >= 1 && iters_run <= p->max_iterations, "iters_run out of range");
  if (ws == nullptr || ws_bytes < plan_workspace(nullptr, 0, N, n, m, sizeof(T)).total) {
    set_error("dicp_icp_backward: workspace too small");
    return DICP_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  IcpWorkspace w = plan_workspace(ws, ws_bytes, N, n, m, sizeof(T));
  const WeightParams P = make_weight_params(p);
  const bool fast = weight_params_fast(P);
  const int nblk = bw_blocks_per_pair(n);
  const size_t rec_stride = (size_t)N * ST_STRIDE;
  const bool pl = p->icp_type == DICP_PT2PL;
  dim3 pgrid(nblk, N);
  const int pair_blocks = (N + 127) / 128;

  DICP_CUDA(cudaMemsetAsync(grad_target, 0, (size_t)N * m * S * sizeof(T), st));
  if (grad_w != nullptr) DICP_CUDA(cudaMemsetAsync(grad_w, 0, (size_t)N * n * sizeof(T), st));
  bwd_seed_kernel<T><<<pgrid, BW_THREADS, 0, st>>>(source, grad_pc, saved_state + (size_t)iters_run * rec_stride, n, grad_source, w.bw_partials);
  DICP_LAUNCH_CHECK();
  for (int k = iters_run - 1; k >= 0; --k) {
    const double* rec_k = saved_state + (size_t)k * rec_stride;
    bwd_pair_kernel<T><<<pair_blocks, 128, 0, st>>>(w.bw_partials, nblk, rec_k, p->dim, k == iters_run - 1 ? 0 : 1, grad_T, w.bw_state, grad_T_init, N);
    DICP_LAUNCH_CHECK();
    const int32_t* idx_k = saved_idx + (size_t)k * N * n;
#define DICP_BW_LAUNCH(PL, FW) \
  bwd_points_kernel<T, PL, FW><<<pgrid, BW_THREADS, 0, st>>>(source, target, w_init, idx_k, rec_k, w.bw_state, n, m, S, P, grad_source, grad_target, grad_w, w.bw_partials)
    if (pl && fast) DICP_BW_LAUNCH(true, true);
    else if (pl) DICP_BW_LAUNCH(true, false);
    else if (fast) DICP_BW_LAUNCH(false, true);
    else DICP_BW_LAUNCH(false, false);
#undef DICP_BW_LAUNCH
    DICP_LAUNCH_CHECK();
  }
  bwd_pair_kernel<T><<<pair_blocks, 128, 0, st>>>(w.bw_partials, nblk, saved_state, p->dim, 2, grad_T, w.bw_state, grad_T_init, N);
  DICP_LAUNCH_CHECK();
  return 0;
}

}  // namespace dicp

extern "C" {
int dicp_icp_backward_f32(const float* source, const float* target, const float* w_init, int N, int n, int m, int S,
                          const dicp_params* params, int iters_run, const int32_t* saved_idx, const double* saved_state,
                          const float* grad_T, const float* grad_pc, float* grad_source, float* grad_target, float* grad_w,
                          float* grad_T_init, void* ws, size_t ws_bytes, void* stream) {
  return dicp::icp_backward_impl<float>(source, target, w_init, N, n, m, S, params, iters_run, saved_idx, saved_state, grad_T, grad_pc,
                                        grad_source, grad_target, grad_w, grad_T_init, ws, ws_bytes, stream);
}
int dicp_icp_backward_f64(const double* source, const double* target, const double* w_init, int N, int n, int m, int S,
                          const dicp_params* params, int iters_run, const int32_t* saved_idx, const double* saved_state,
                          const double* grad_T, const double* grad_pc, double* grad_source, double* grad_target, double* grad_w,
                          double* grad_T_init, void* ws, size_t ws_bytes, void* stream) {
  return dicp::icp_backward_impl<double>(source, target, w_init, N, n, m, S, params, iters_run, saved_idx, saved_state, grad_T, grad_pc,
                                         grad_source, grad_target, grad_w, grad_T_init, ws, ws_bytes, stream);
}
}
