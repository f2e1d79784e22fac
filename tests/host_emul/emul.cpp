// emul.cpp — TEST INFRASTRUCTURE ONLY.  Compiles the __host__ __device__ arithmetic of
// dicp_b200/csrc/dicp_math.cuh with g++ and drives it with plain sequential loops that
// mirror the CUDA launch sequence (nn scan -> gn accumulate -> pair update; reverse:
// pair backward -> point backward).  This lets the GPU-less CI check the per-point math,
// the 6x6 solve, the exp-map update, the bookkeeping and the adjoint against the oracle and
// the golden vectors.  The product (dicp_b200/) never loads this library.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/dicp_b200.h"
#include "../../dicp_b200/csrc/dicp_math.cuh"

using namespace dicp;

static WeightParams wp_of(const dicp_params* p) {
  WeightParams w;
  w.loss_id = p->loss_id; w.use_trim = p->use_trim; w.diff = p->differentiable;
  w.c = p->loss_metric; w.tau = p->trim_dist; w.steep = p->tanh_steepness; w.thresh = p->match_ratio_thresh;
  return w;
}

template <typename T, bool PL, int DIM>
static void accum_pair(const T* src, const T* tgt, const T* w_init, const int32_t* idx, const double* rec, int n, int m, int S,
                       const WeightParams& P, T* wout, double* red) {
  T C[9], r[3];
  for (int i = 0; i < 9; ++i) C[i] = (T)rec[ST_C + i];
  for (int i = 0; i < 3; ++i) r[i] = (T)rec[ST_R + i];
  const T active = (T)rec[ST_ACTIVE];
  // same summation structure as the kernel: a few strided per-"thread" partial sums in T, then double
  const int lanes = 64;
  for (int v = 0; v < RED_N; ++v) red[v] = 0.0;
  for (int l = 0; l < lanes; ++l) {
    T acc[RED_N];
    for (int v = 0; v < RED_N; ++v) acc[v] = T(0);
    for (int i = l; i < n; i += lanes) {
      const T* p = src + 3 * (size_t)i;
      const T* row = tgt + (size_t)idx[i] * S;
      T nr[3] = {T(0), T(0), T(0)};
      if (PL) { nr[0] = row[3]; nr[1] = row[4]; nr[2] = row[5]; }
      const T w0 = (w_init ? w_init[i] : T(1)) * active;
      PointEval<T> s;
      eval_point<T, PL>(C, r, p, row, nr, w0, P, s);
      accum_point<T, PL, DIM>(acc, s, w0, (T)P.thresh);
      const int R = PL ? 1 : 3;
      if (wout) for (int k = 0; k < R; ++k) wout[(size_t)i * R + k] = s.w;
    }
    for (int v = 0; v < RED_N; ++v) red[v] += (double)acc[v];
  }
}

template <typename T>
static int forward(const T* source, const T* target, const T* T_init, const T* w_init, int N, int n, int m, int S,
                   const dicp_params* p, T* pc, T* T_out, T* deltas, T* weights, T* costs, uint8_t* converged, T* num_iters,
                   T* match_ratio, int32_t* saved_idx, double* saved_state, int32_t* iters_run) {
  const WeightParams P = wp_of(p);
  const int K = p->max_iterations;
  const bool pl = p->icp_type == DICP_PT2PL;
  const long long nw = (long long)n * (pl ? 1 : 3);
  const size_t rs = (size_t)N * ST_STRIDE;
  memset(saved_state, 0, sizeof(double) * rs);
  for (int b = 0; b < N; ++b) {
    double* rec = saved_state + (size_t)b * ST_STRIDE;
    for (int rr = 0; rr < 3; ++rr) {
      for (int c = 0; c < 3; ++c) rec[ST_C + rr * 3 + c] = (double)T_init[b * 16 + rr * 4 + c];
      rec[ST_R + rr] = (double)T_init[b * 16 + rr * 4 + 3];
    }
    rec[ST_ACTIVE] = 1.0;
    converged[b] = 0; num_iters[b] = T(0); match_ratio[b] = T(0);
  }
  int k_run = 0;
  bool singular = false;
  for (int k = 0; k < K; ++k) {
    double* rec_k = saved_state + (size_t)k * rs;
    int nbelow = 0;
    for (int b = 0; b < N; ++b) {
      double* cur = rec_k + (size_t)b * ST_STRIDE;
      double* nxt = cur + rs;
      const T* src = source + (size_t)b * n * 3;
      const T* tgt = target + (size_t)b * m * S;
      int32_t* idx = saved_idx + ((size_t)k * N + b) * n;
      T C[9], r[3];
      for (int i = 0; i < 9; ++i) C[i] = (T)cur[ST_C + i];
      for (int i = 0; i < 3; ++i) r[i] = (T)cur[ST_R + i];
      for (int i = 0; i < n; ++i) {                         // brute-force NN, direct form, first minimum
        T a[3];
        rotate(C, src + 3 * (size_t)i, a);
        const T x0 = a[0] + r[0], x1 = a[1] + r[1], x2 = a[2] + r[2];
        T best = 0; int bi = -1;
        for (int j = 0; j < m; ++j) {
          const T d = dist2<T>(x0, x1, x2, tgt[(size_t)j * S], tgt[(size_t)j * S + 1], tgt[(size_t)j * S + 2]);
          if (bi < 0 || d < best) { best = d; bi = j; }
        }
        idx[i] = bi;
      }
      double red[RED_N];
      T* wout = weights ? weights + ((size_t)b * K + k) * nw : nullptr;
      const T* w0 = w_init ? w_init + (size_t)b * n : nullptr;
      if (pl && p->dim == 3) accum_pair<T, true, 3>(src, tgt, w0, idx, cur, n, m, S, P, wout, red);
      else if (pl) accum_pair<T, true, 2>(src, tgt, w0, idx, cur, n, m, S, P, wout, red);
      else if (p->dim == 3) accum_pair<T, false, 3>(src, tgt, w0, idx, cur, n, m, S, P, wout, red);
      else accum_pair<T, false, 2>(src, tgt, w0, idx, cur, n, m, S, P, wout, red);
      for (int i = 0; i < ST_STRIDE; ++i) nxt[i] = 0.0;
      double ni = (double)num_iters[b], mr = (double)match_ratio[b];
      const double prev_cost = k > 0 ? (double)costs[(size_t)b * K + k - 1] : 0.0;
      PairOutcome out;
      pair_update<T>(red, cur, nxt, k, p->dim, p->const_iter != 0, p->tolerance, prev_cost, &ni, &mr, out);
      for (int i = 0; i < 6; ++i) deltas[((size_t)b * K + k) * 6 + i] = (T)out.delta[i];
      costs[(size_t)b * K + k] = (T)out.cost;
      num_iters[b] = (T)ni; match_ratio[b] = (T)mr;
      if (out.below) { converged[b] = 1; ++nbelow; }
      if (out.singular) singular = true;
      if (out.dead_weights && wout) for (long long i = 0; i < nw; ++i) wout[i] = wout[i - nw];
    }
    k_run = k + 1;
    if (!p->const_iter && nbelow == N) break;
  }
  for (int b = 0; b < N; ++b) {
    const double* after = saved_state + (size_t)k_run * rs + (size_t)b * ST_STRIDE;
    const double* last = saved_state + (size_t)(k_run - 1) * rs + (size_t)b * ST_STRIDE;
    T C[9], r[3];
    for (int i = 0; i < 9; ++i) C[i] = (T)after[ST_C + i];
    for (int i = 0; i < 3; ++i) r[i] = (T)after[ST_R + i];
    for (int i = 0; i < n; ++i) {
      T a[3];
      rotate(C, source + ((size_t)b * n + i) * 3, a);
      for (int c = 0; c < 3; ++c) pc[((size_t)b * n + i) * 3 + c] = a[c] + r[c];
    }
    T* To = T_out + (size_t)b * 16;
    for (int rr = 0; rr < 3; ++rr) { for (int c = 0; c < 3; ++c) To[rr * 4 + c] = C[rr * 3 + c]; To[rr * 4 + 3] = r[rr]; }
    To[12] = To[13] = To[14] = T(0); To[15] = T(1);
    double ni = (double)num_iters[b], mr = (double)match_ratio[b];
    final_stats(last, after, k_run, p->match_ratio_thresh, (double)nw, &ni, &mr);
    num_iters[b] = (T)ni; match_ratio[b] = (T)mr;
  }
  *iters_run = k_run;
  return singular ? DICP_ERR_SINGULAR : 0;
}

template <typename T, bool PL>
static void backward_points(const T* src, const T* tgt, const T* w_init, const int32_t* idx, const double* rec, const double* lamd,
                            int n, int S, const WeightParams& P, T* gs, T* gt, T* gw, double* carry) {
  T C[9], r[3], lam[6], del[6];
  for (int i = 0; i < 9; ++i) C[i] = (T)rec[ST_C + i];
  for (int i = 0; i < 3; ++i) r[i] = (T)rec[ST_R + i];
  for (int i = 0; i < 6; ++i) { lam[i] = (T)lamd[i]; del[i] = (T)rec[ST_DELTA + i]; }
  for (int i = 0; i < n; ++i) {
    const T* p = src + 3 * (size_t)i;
    const T* row = tgt + (size_t)idx[i] * S;
    T nr[3] = {T(0), T(0), T(0)};
    if (PL) { nr[0] = row[3]; nr[1] = row[4]; nr[2] = row[5]; }
    const T w0 = w_init ? w_init[i] : T(1);
    PointEval<T> s;
    eval_point<T, PL>(C, r, p, row, nr, w0, P, s);
    T abar[3], ebar[3], nbar[3], w0bar;
    point_backward<T, PL>(s, w0, lam, del, abar, ebar, nbar, w0bar);
    if (gw) gw[i] += w0bar;
    for (int c = 0; c < 3; ++c) gs[3 * (size_t)i + c] += C[c] * abar[0] + C[3 + c] * abar[1] + C[6 + c] * abar[2];
    T* grow = gt + (size_t)idx[i] * S;
    for (int c = 0; c < 3; ++c) grow[c] += -ebar[c];
    if (PL) for (int c = 0; c < 3; ++c) grow[3 + c] += nbar[c];
    for (int rr = 0; rr < 3; ++rr) {
      for (int c = 0; c < 3; ++c) carry[rr * 3 + c] += (double)(abar[rr] * p[c]);
      carry[9 + rr] += (double)ebar[rr];
    }
  }
}

template <typename T>
static int backward(const T* source, const T* target, const T* w_init, int N, int n, int m, int S, const dicp_params* p, int iters_run,
                    const int32_t* saved_idx, const double* saved_state, const T* grad_T, const T* grad_pc, T* g_src, T* g_tgt,
                    T* g_w, T* g_T0) {
  const WeightParams P = wp_of(p);
  const size_t rs = (size_t)N * ST_STRIDE;
  const bool pl = p->icp_type == DICP_PT2PL;
  memset(g_tgt, 0, sizeof(T) * (size_t)N * m * S);
  if (g_w) memset(g_w, 0, sizeof(T) * (size_t)N * n);
  for (int b = 0; b < N; ++b) {
    const T* src = source + (size_t)b * n * 3;
    T* gs = g_src + (size_t)b * n * 3;
    double carry[12];
    for (int rr = 0; rr < 3; ++rr) {
      for (int c = 0; c < 3; ++c) carry[rr * 3 + c] = grad_T ? (double)grad_T[b * 16 + rr * 4 + c] : 0.0;
      carry[9 + rr] = grad_T ? (double)grad_T[b * 16 + rr * 4 + 3] : 0.0;
    }
    const double* fin = saved_state + (size_t)iters_run * rs + (size_t)b * ST_STRIDE;
    T C[9];
    for (int i = 0; i < 9; ++i) C[i] = (T)fin[ST_C + i];
    for (int i = 0; i < n; ++i) {
      T g[3] = {T(0), T(0), T(0)};
      if (grad_pc) for (int c = 0; c < 3; ++c) g[c] = grad_pc[((size_t)b * n + i) * 3 + c];
      for (int c = 0; c < 3; ++c) gs[3 * i + c] = C[c] * g[0] + C[3 + c] * g[1] + C[6 + c] * g[2];
      for (int rr = 0; rr < 3; ++rr) {
        for (int c = 0; c < 3; ++c) carry[rr * 3 + c] += (double)(g[rr] * src[3 * i + c]);
        carry[9 + rr] += (double)g[rr];
      }
    }
    for (int k = iters_run - 1; k >= 0; --k) {
      const double* rec = saved_state + (size_t)k * rs + (size_t)b * ST_STRIDE;
      if (rec[ST_ACTIVE] == 0.0) continue;
      double lam[6];
      pair_backward(rec, p->dim, carry, carry + 9, lam);
      const int32_t* idx = saved_idx + ((size_t)k * N + b) * n;
      const T* w0 = w_init ? w_init + (size_t)b * n : nullptr;
      T* gw = g_w ? g_w + (size_t)b * n : nullptr;
      if (pl) backward_points<T, true>(src, target + (size_t)b * m * S, w0, idx, rec, lam, n, S, P, gs, g_tgt + (size_t)b * m * S, gw, carry);
      else backward_points<T, false>(src, target + (size_t)b * m * S, w0, idx, rec, lam, n, S, P, gs, g_tgt + (size_t)b * m * S, gw, carry);
    }
    T* g = g_T0 + (size_t)b * 16;
    for (int rr = 0; rr < 3; ++rr) { for (int c = 0; c < 3; ++c) g[rr * 4 + c] = (T)carry[rr * 3 + c]; g[rr * 4 + 3] = (T)carry[9 + rr]; }
    g[12] = g[13] = g[14] = g[15] = T(0);
  }
  return 0;
}

extern "C" {
#define EMUL_FWD(SFX, T)                                                                                                          \
  __attribute__((visibility("default"))) int emul_icp_forward_##SFX(                                                              \
      const T* source, const T* target, const T* T_init, const T* w_init, const int32_t*, const int32_t*, int N, int n, int m, int S, \
      const dicp_params* p, T* pc,                                                                                               \
      T* T_out, T* deltas, T* weights, T* costs, uint8_t* converged, T* num_iters, T* match_ratio, int32_t* saved_idx,            \
      double* saved_state, int32_t* iters_run, void*, size_t, void*) {                                                            \
    return forward<T>(source, target, T_init, w_init, N, n, m, S, p, pc, T_out, deltas, weights, costs, converged, num_iters,      \
                      match_ratio, saved_idx, saved_state, iters_run);                                                            \
  }                                                                                                                               \
  __attribute__((visibility("default"))) int emul_icp_backward_##SFX(                                                             \
      const T* source, const T* target, const T* w_init, const int32_t*, int N, int n, int m, int S, const dicp_params* p, int iters_run, \
      const int32_t* saved_idx, const double* saved_state, const T* grad_T, const T* grad_pc, T* g_src, T* g_tgt, T* g_w,         \
      T* g_T0, void*, size_t, void*) {                                                                                            \
    return backward<T>(source, target, w_init, N, n, m, S, p, iters_run, saved_idx, saved_state, grad_T, grad_pc, g_src, g_tgt,   \
                       g_w, g_T0);                                                                                                \
  }
EMUL_FWD(f32, float)
EMUL_FWD(f64, double)
// Kabsch / SVD step (dicp_math.cuh: svd3, kabsch_step, kabsch_step_backward)
__attribute__((visibility("default"))) void emul_kabsch_step(const double* sums, double n, int literal, double* rec) {
  dicp::kabsch_step(sums, n, literal != 0, rec);
}
__attribute__((visibility("default"))) void emul_kabsch_step_backward(const double* rec, double n, int literal, const double* Cbar,
                                                                      const double* rbar, double* g) {
  dicp::kabsch_step_backward(rec, n, literal != 0, Cbar, rbar, g);
}
}
