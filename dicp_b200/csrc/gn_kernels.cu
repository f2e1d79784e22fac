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
#include "tile_stream.cuh"

namespace dicp {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

#ifndef GN_NT
#define GN_NT 128
#endif
constexpr int GN_THREADS = GN_NT;
#ifndef GN_ROWS24
#define GN_ROWS24 0       // 1: point-to-plane gathers the caller's 24-byte rows directly (two aligned 16-byte loads) instead of the 32-byte
                          // table: 15 % less DRAM traffic (184 vs 215 MB read at C3) but 1.5 sectors per row through L2: 71 vs 53 us, rejected
#endif
#ifndef GN_U
#define GN_U 2
#endif
#ifndef GN_STAGES
#define GN_STAGES 4
#endif
#ifndef GN_MINB
#define GN_MINB 4
#endif
constexpr int GN_TILE = GN_THREADS * GN_U;      // correspondences per tile (one staged pass of the CTA)
constexpr int GN_CTAS_PER_SM = GN_MINB;         // persistent CTAs per SM (the launch bounds guarantee the registers)

int device_sm_count() {
  static int cached = 0;                        // benign race: every thread computes the same value
  if (cached == 0) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) {
      cudaGetLastError();
      sms = 148;
    }
    cached = sms;
  }
  return cached;
}

#ifndef GN_K
#define GN_K 8
#endif
#ifndef GRID_WARM
#define GRID_WARM 0        // 1: the grid search of iteration k > 0 starts from the neighbours of iteration k - 1 (a real candidate bounds the
                           // search from the start; same result).  Measured on top of the row pruning of grid_query_kernel: 5.14 ms per C3
                           // step instead of 5.01 -- the extra gather costs more than the tighter first bound saves; off
#endif
#ifndef GN_L2PF
#define GN_L2PF 0
#endif
#ifndef GN_L2HINT
#define GN_L2HINT 0
#endif
TilePlan gn_tile_plan(int N, int n) { return make_tile_plan(N, n, GN_TILE, GN_CTAS_PER_SM, GN_K); }
int gn_blocks_per_pair(int N, int n) { return gn_tile_plan(N, n).rows; }


// block-wide sum of NV (<= 32) per-thread values: per warp a 31-shuffle multi-value reduction in the compute type
// (tile_stream.cuh), then double across warps; result lands in out[0..NV) of the segment's partial row.
// Fixed tree => deterministic for a given launch shape.
template <typename T, int NV>
__device__ __forceinline__ void block_reduce_store(const T* vals, double* out, double* smem /* [warps][32] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  T v[32];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = vals[i];
  warp_multi_reduce<T, NV>(v);
  smem[warp * 32 + lane] = (double)v[0];
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0.0;
    for (int w = 0; w < nwarps; ++w) s += smem[w * 32 + threadIdx.x];
    out[threadIdx.x] = s;
  }
  __syncthreads();                               // smem is reused by the next segment this CTA reduces
}

// What a persistent CTA keeps per (chunk, pair) segment it is working on, and the hand-over between segments: the
// partial sums go to the row of the pair's RED_N-wide table that the plan assigns to the segment (the number of
// chunk boundaries between the start of the pair and the start of the segment); the segment that ends the pair also
// clears the rows nobody writes, so that the update kernel can sum all plan.rows rows in index order.
template <typename T>
struct GnPairCtx {
  T C[9], r[3], active;
  int b;
  long long seg_tile;     // first tile of the segment
};
template <typename T>
__device__ __forceinline__ void gn_flush_segment(const TilePlan& plan, int b, long long seg_tile, long long last_tile, const T* acc,
                                                 double* __restrict__ partials, double* red_smem) {
  const long long pair_tile = (long long)b * plan.tpp;
  const int row = tp_segment_row(plan, b, seg_tile);
  double* rows_b = partials + (size_t)b * plan.rows * RED_N;
  block_reduce_store<T, RED_N>(acc, rows_b + (size_t)row * RED_N, red_smem);
  if (last_tile == pair_tile + plan.tpp - 1) {
    for (int i = (row + 1) * RED_N + threadIdx.x; i < plan.rows * RED_N; i += blockDim.x) rows_b[i] = 0.0;
  }
}
template <typename T>
__device__ __forceinline__ void gn_enter_segment(GnPairCtx<T>& ctx, int b, long long seg_tile, const StateView& state, T* acc) {
  if (b != ctx.b) {
    const double* rec = state.base + (size_t)b * state.stride;
#pragma unroll
    for (int i = 0; i < 9; ++i) ctx.C[i] = (T)rec[state.c + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) ctx.r[i] = (T)rec[state.r + i];
    ctx.active = state.active >= 0 ? (T)rec[state.active] : T(1);
    ctx.b = b;
  }
  ctx.seg_tile = seg_tile;
#pragma unroll
  for (int i = 0; i < RED_N; ++i) acc[i] = T(0);
}

// ---------------------------------------------------------------------------------------------------------
// fp32 fast path: streams staged in shared memory by TMA bulk copies, target rows gathered one tile ahead.
// Requirements (checked by the launcher): 16-byte aligned source / w_init and `rows` built.  With n % 4 == 0 (and an
// aligned idx) every slice of each stream starts on a 16-byte boundary and has a 16-byte multiple length; otherwise
// the UNAL instantiation copies the aligned superset of every slice (see st_off / issue).
//
// The CTA's ring has STAGES stages of { idx[TILE] | p[3 TILE] | w0[TILE] } with a "full" mbarrier (bulk-copy
// transaction bytes) and an "empty" mbarrier (one arrival per warp) per stage.  No __syncthreads separates
// tiles: every warp waits for a stage on its own, releases it on its own, and the warp whose turn it is (tile L
// is fed by warp L % WARPS) refills the stage that was released one step earlier, so the warps of a CTA drift by up
// to a tile and their memory waits and arithmetic interleave instead of lining up (16 loosely coupled streams
// per SM instead of 4 lock-stepped CTAs).  Lane l of warp w owns correspondences [U (32 w + l), +U) of the tile
// (U = 2: 256-correspondence tiles): its indices / prior weights / weights row are one 8-byte access each, its points
// three.  The (C, r, active) record of the next segment's pair is fetched one step ahead with cp.async into a
// per-warp slot.  The rows are gathered with ONE request each (LDG.256 on the aligned table): every variant with two
// requests per row (24-byte rows, cp.async into shared memory) measured slower -- the gather is bound by the request
// rate through L1 / L2, not by registers or DRAM bytes.
// ---------------------------------------------------------------------------------------------------------
template <bool PT2PL, int DIM, bool FASTW, bool RAGGED, bool UNAL>
__global__ void __launch_bounds__(GN_THREADS, GN_MINB)
gn_accum_staged_kernel(const float* __restrict__ source, const float* __restrict__ target, const void* __restrict__ rows,
                       const float* __restrict__ w_init, const int32_t* __restrict__ idx, StateView state, int n, int m, WeightParams P,
                       float* __restrict__ weights_out, long long w_pair_stride, double* __restrict__ partials, TilePlan plan,
                       const int* __restrict__ stop_flag, int N, const int32_t* __restrict__ n_valid) {
  if (stop_flag != nullptr && *stop_flag != 0) return;
  constexpr int U = GN_U, TILE = GN_TILE, STAGES = GN_STAGES, R = PT2PL ? 1 : 3, WARPS = GN_THREADS / 32;
  static_assert(STAGES >= 3, "a stage is refilled one step after its release");
  // dynamic shared memory: STAGES x { idx[TILE] | p[3 TILE] | w0[TILE] } (20 TILE bytes), reduction scratch, state slots, barriers
  extern __shared__ __align__(128) unsigned char gn_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // UNAL: the clouds' length is not a multiple of 4, so a tile's slice of the streams does not start on a 16-byte
  // boundary.  The feeder then copies the aligned superset of the slice (up to three elements more at either end; a
  // stage has room for TILE + 8) and the consumers read at the slice's offset inside it.
  constexpr int STG = TILE + (UNAL ? 8 : 0);
  auto st_idx = [&](int s) { return reinterpret_cast<int32_t*>(gn_smem + (size_t)s * STG * 20); };
  auto st_p = [&](int s) { return reinterpret_cast<float*>(gn_smem + (size_t)s * STG * 20 + STG * 4); };
  auto st_w = [&](int s) { return reinterpret_cast<float*>(gn_smem + (size_t)s * STG * 20 + STG * 16); };
  auto st_off = [&](const TileCursor& c) { return UNAL ? ((c.b & 3) * (n & 3)) & 3 : 0; };      // (b n + j TILE) mod 4
  const int idx_mis = UNAL ? (int)((reinterpret_cast<uintptr_t>(idx) >> 2) & 3) : 0;              // elements past a 16-byte boundary
  double* red_smem = reinterpret_cast<double*>(gn_smem + (size_t)STAGES * STG * 20);
  double* slot = red_smem + WARPS * 32 + warp * 16;            // this warp's (C, r, active) record of the next segment
  uint64_t* full = reinterpret_cast<uint64_t*>(red_smem + WARPS * 32 + WARPS * 16);
  uint64_t* empty = full + STAGES;
  TileWalk walk;
  walk.init(plan, blockIdx.x);
  const int ntiles = walk.ntiles;
  if (ntiles == 0) return;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], WARPS); }
    mbar_fence_init();
  }
  __syncthreads();

  // Tile L of this CTA lives in ring stage L % STAGES, use number L / STAGES of that stage.
  typedef TileCursor Cursor;
  auto tile_count = [&](const Cursor& c) { return min(TILE, n - c.j * TILE); };
  // correspondences of the tile that are real: rows >= n_valid[b] of a ragged batch are padding (never evaluated, weight 0)
  // (RAGGED is a compile-time option: the per-tile look-up and the masks cost the dense kernel 3 %)
  auto tile_valid = [&](const Cursor& c) {
    if (!RAGGED) return min(TILE, n - c.j * TILE);
    return max(0, min(TILE, min(n_valid[c.b], n) - c.j * TILE));
  };

  // one lane feeds the ring
  auto issue = [&](const Cursor& c) {
    const int s = c.L % STAGES;
    const int i0 = c.j * TILE;
    const uint32_t cnt = (uint32_t)tile_count(c);
    const size_t o = (size_t)c.b * n + i0;
    if (UNAL) {
      // aligned superset [a0, a0 + nb) of the slice [o, o + cnt) of a stream whose base is `mis` elements past a 16-byte
      // boundary; it never passes the last whole 16-byte group of the array (what is left of the slice behind it, at
      // most three elements, is copied with ordinary loads and stores and released by the arrival below)
      const long long total = (long long)N * n, oo = (long long)o;
      auto span = [&](int mis, long long& a0, uint32_t& nb) {
        a0 = ((oo + mis) & ~3LL) - mis;
        long long a1 = ((oo + cnt + mis + 3) & ~3LL) - mis;
        const long long lim = ((total + mis) & ~3LL) - mis;
        if (a1 > lim) a1 = lim;
        nb = a1 > a0 ? (uint32_t)(a1 - a0) : 0u;
      };
      long long a0, ai0;
      uint32_t nb, nbi;
      span(0, a0, nb);                 // source, w_init: 16-byte aligned bases
      span(idx_mis, ai0, nbi);         // the indices of iteration k start k N n elements into the saved buffer
      for (long long e = (a0 + nb > oo ? a0 + nb : oo); e < oo + cnt; ++e) {
        st_p(s)[(e - a0) * 3] = source[e * 3]; st_p(s)[(e - a0) * 3 + 1] = source[e * 3 + 1]; st_p(s)[(e - a0) * 3 + 2] = source[e * 3 + 2];
        if (w_init != nullptr) st_w(s)[e - a0] = w_init[e];
      }
      for (long long e = (ai0 + nbi > oo ? ai0 + nbi : oo); e < oo + cnt; ++e) st_idx(s)[e - ai0] = idx[e];
      mbar_expect_tx(&full[s], nbi * 4u + nb * (w_init != nullptr ? 16u : 12u));
      if (nbi > 0u) bulk_g2s(st_idx(s), idx + ai0, nbi * 4u, &full[s]);
      if (nb > 0u) {
        bulk_g2s(st_p(s), source + a0 * 3, nb * 12u, &full[s]);
        if (w_init != nullptr) bulk_g2s(st_w(s), w_init + a0, nb * 4u, &full[s]);
      }
      return;
    }
    mbar_expect_tx(&full[s], cnt * (w_init != nullptr ? 20u : 16u));
#if GN_L2HINT
    const uint64_t pol_stream = l2_policy_evict_first();       // read-once streams must not push the gather table out of L2
    bulk_g2s_hint(st_idx(s), idx + o, cnt * 4u, &full[s], pol_stream);
    bulk_g2s_hint(st_p(s), source + o * 3, cnt * 12u, &full[s], pol_stream);
    if (w_init != nullptr) bulk_g2s_hint(st_w(s), w_init + o, cnt * 4u, &full[s], pol_stream);
#else
    bulk_g2s(st_idx(s), idx + o, cnt * 4u, &full[s]);
    bulk_g2s(st_p(s), source + o * 3, cnt * 12u, &full[s]);
    if (w_init != nullptr) bulk_g2s(st_w(s), w_init + o, cnt * 4u, &full[s]);
#endif
#if GN_L2PF
    // Experiment (off): stream the gather table into L2 one wave ahead of the random row gathers with
    // cp.async.bulk.prefetch.L2 -- the tile at position (b, j) pulls the j-th slice of the table of the pair that sits
    // G k tiles further on (every tile of the next wave is named exactly once this way), the tiles of a CTA's first
    // chunk pull their own pair's slices -- so that DRAM sees sequential 8 KB bursts and the gathers hit L2.  Measured:
    // 60.3 instead of 53.7 us at C3, DRAM reads 263 instead of 215 MB (the prefetched lines are evicted by the 100 MB of
    // streams that pass through L2 before they are used, or fetched twice); with evict_last on the prefetch and
    // evict_first on the streams (GN_L2HINT) 59.3 us; the hints alone change nothing (53.7).
    if (rows != nullptr) {
      const int rec_bytes = PT2PL ? 32 : 16;
      const int rpt = (m + plan.tpp - 1) / plan.tpp;
      auto pull = [&](int pb, int pj) {
        const int r0 = pj * rpt, r1 = min(m, r0 + rpt);
#if GN_L2HINT
        if (r0 < r1) bulk_prefetch_l2_hint(static_cast<const char*>(rows) + ((size_t)pb * m + r0) * rec_bytes, (uint32_t)(r1 - r0) * rec_bytes, l2_policy_evict_last());
#else
        if (r0 < r1) bulk_prefetch_l2(static_cast<const char*>(rows) + ((size_t)pb * m + r0) * rec_bytes, (uint32_t)(r1 - r0) * rec_bytes);
#endif
      };
      if (c.wave == 0) pull(c.b, c.j);
      int b2 = c.b + plan.adv_b, j2 = c.j + plan.adv_j;
      if (j2 >= plan.tpp) { j2 -= plan.tpp; ++b2; }
      if ((long long)b2 * plan.tpp + j2 < plan.tiles) pull(b2, j2);
    }
#endif
  };
  // asynchronous fetch of a pair's (C, r, active) doubles into the warp's slot
  auto prefetch_state = [&](int b) {
    const double* rec = state.base + (size_t)b * state.stride;
    if (lane < 12) cp_async8(slot + lane, rec + (lane < 9 ? state.c + lane : state.r + lane - 9));
    else if (lane == 12 && state.active >= 0) cp_async8(slot + 12, rec + state.active);
  };

  struct Gathered { float q[U][3], nr[U][3]; };
  // waits for the tile's streams, then issues the row gathers of this lane's U correspondences
  auto gather = [&](const Cursor& c, Gathered& g) {
    const int s = c.L % STAGES;
    const int cnt = tile_valid(c);
    mbar_wait(&full[s], (uint32_t)((c.L / STAGES) & 1));
    const size_t row_b = (size_t)c.b * m;
    const bool last_pair = c.b == N - 1;                       // the very last row of the tensor must not be over-read
    int jj[U];
    const int off = st_off(c);
    if (U == 4 && !UNAL) {
      const int4 v = *reinterpret_cast<const int4*>(st_idx(s) + tid * U);
      jj[0] = v.x; jj[1 % U] = v.y; jj[2 % U] = v.z; jj[3 % U] = v.w;
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u) jj[u] = st_idx(s)[((off + idx_mis) & 3) + tid * U + u];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = tid * U + u < cnt ? jj[u] : 0;           // lanes past the end of the cloud gather row 0 and are masked below
      if (PT2PL && GN_ROWS24) load_row24(target + (row_b + j) * 6, last_pair && j == m - 1, g.q[u], g.nr[u]);
      else load_packed_row<float, PT2PL>(rows, row_b + j, g.q[u], g.nr[u]);
    }
  };

  float acc[RED_N];
  GnPairCtx<float> ctx;
  ctx.b = -1;
  long long last_tile = 0;
  struct Staged { float p[U][3], w0[U]; };
  // this lane's points and prior weights of the tile, from the ring
  auto read_stage = [&](const Cursor& c, Staged& in) {
    const int s = c.L % STAGES;
    const int off = st_off(c);
    const float4* sp = reinterpret_cast<const float4*>(st_p(s) + tid * U * 3);
    float flat[U * 3];
    if (U == 4 && !UNAL) {
      const float4 a = sp[0], bq = sp[1], cq = sp[2];
      flat[0] = a.x; flat[1] = a.y; flat[2] = a.z; flat[3] = a.w;
      flat[4 % (U * 3)] = bq.x; flat[5 % (U * 3)] = bq.y; flat[6 % (U * 3)] = bq.z; flat[7 % (U * 3)] = bq.w;
      flat[8 % (U * 3)] = cq.x; flat[9 % (U * 3)] = cq.y; flat[10 % (U * 3)] = cq.z; flat[11 % (U * 3)] = cq.w;
    } else {
#pragma unroll
      for (int k = 0; k < U * 3; ++k) flat[k] = st_p(s)[(off + tid * U) * 3 + k];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) { in.p[u][0] = flat[3 * u]; in.p[u][1] = flat[3 * u + 1]; in.p[u][2] = flat[3 * u + 2]; }
    if (w_init != nullptr) {
      if (U == 4 && !UNAL) {
        const float4 w4 = *reinterpret_cast<const float4*>(st_w(s) + tid * U);
        in.w0[0] = w4.x; in.w0[1 % U] = w4.y; in.w0[2 % U] = w4.z; in.w0[3 % U] = w4.w;
      } else {
#pragma unroll
        for (int u = 0; u < U; ++u) in.w0[u] = st_w(s)[off + tid * U + u];
      }
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u) in.w0[u] = 1.0f;
    }
  };
  auto compute = [&](const Cursor& c, const Gathered& g, Staged& in) {
    const int b = c.b;
    const int i0 = c.j * TILE;
    const int cnt = tile_valid(c), cnt_all = tile_count(c);
#pragma unroll
    for (int u = 0; u < U; ++u) in.w0[u] *= ctx.active;                                           // ICP.py:257
    PointEval<float> ev[U];
#pragma unroll
    for (int u = 0; u < U; ++u) eval_point<float, PT2PL, FASTW>(ctx.C, ctx.r, in.p[u], g.q[u], g.nr[u], in.w0[u], P, ev[u]);
    if (cnt == TILE) {
#pragma unroll
      for (int u = 0; u < U; ++u) accum_point<float, PT2PL, DIM>(acc, ev[u], in.w0[u], (float)P.thresh);
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (tid * U + u < cnt) accum_point<float, PT2PL, DIM>(acc, ev[u], in.w0[u], (float)P.thresh);
    }
    if (weights_out != nullptr) {                                                                 // ICP.py:222, 164-165
      float* wrow = weights_out + (size_t)b * w_pair_stride + (size_t)(i0 + tid * U) * R;
      float wv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) wv[u] = tid * U + u < cnt ? ev[u].w : 0.0f;        // padding rows of a ragged batch carry weight 0
      if (U == 4 && cnt_all == TILE && (reinterpret_cast<uintptr_t>(wrow) & 15) == 0) {
        if (R == 1) {
          __stcs(reinterpret_cast<float4*>(wrow), make_float4(wv[0], wv[1 % U], wv[2 % U], wv[3 % U]));
        } else {
          __stcs(reinterpret_cast<float4*>(wrow), make_float4(wv[0], wv[0], wv[0], wv[1 % U]));
          __stcs(reinterpret_cast<float4*>(wrow) + 1, make_float4(wv[1 % U], wv[1 % U], wv[2 % U], wv[2 % U]));
          __stcs(reinterpret_cast<float4*>(wrow) + 2, make_float4(wv[2 % U], wv[3 % U], wv[3 % U], wv[3 % U]));
        }
      } else if (U == 2 && cnt_all == TILE && (reinterpret_cast<uintptr_t>(wrow) & 7) == 0) {
        if (R == 1) {
          __stcs(reinterpret_cast<float2*>(wrow), make_float2(wv[0], wv[1 % U]));
        } else {
          __stcs(reinterpret_cast<float2*>(wrow), make_float2(wv[0], wv[0]));
          __stcs(reinterpret_cast<float2*>(wrow) + 1, make_float2(wv[0], wv[1 % U]));
          __stcs(reinterpret_cast<float2*>(wrow) + 2, make_float2(wv[1 % U], wv[1 % U]));
        }
      } else {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (tid * U + u >= cnt_all) continue;
#pragma unroll
          for (int k = 0; k < R; ++k) wrow[u * R + k] = wv[u];
        }
      }
    }
  };
  // start of a (chunk, pair) segment: hand over the sums of the previous one, take the pair's record from the slot
  auto enter = [&](const Cursor& c) {
    if (c.L > 0) gn_flush_segment<float>(plan, ctx.b, ctx.seg_tile, last_tile, acc, partials, red_smem);
    cp_async_wait_all();
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 9; ++i) ctx.C[i] = (float)slot[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) ctx.r[i] = (float)slot[9 + i];
    ctx.active = state.active >= 0 ? (float)slot[12] : 1.0f;
    ctx.b = c.b;
    ctx.seg_tile = c.tile(plan);
#pragma unroll
    for (int i = 0; i < RED_N; ++i) acc[i] = 0.0f;
    __syncwarp();                                              // the slot may be refilled from here on
  };

  Cursor ci, cg;                                  // next tile to feed into the ring / tile whose rows are being gathered
  ci.init(plan, walk);
  cg.init(plan, walk);
  // one pipeline step: tile `cc` is computed with the rows gathered one step earlier (gc) while the rows of the next
  // tile (cg -> gn) are requested; shared-memory reads of this tile go first so that they do not queue behind the gathers
  auto step = [&](const Gathered& gc, Gathered& gn) {
    const Cursor cc = cg;
    if (cc.seg_first) enter(cc);                               // CTA-uniform: every warp walks the same tiles
    last_tile = cc.tile(plan);
    Staged in;
    read_stage(cc, in);
    cg.next(plan, walk);
    if (cg.L < ntiles) {
      if (cg.seg_first) prefetch_state(cg.b);
      gather(cg, gn);
    }
    // this warp is done with the stage of tile cc (indices at gather time, points / weights just now)
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[cc.L % STAGES]);
    // the ring is fed one step behind the releases: tile ci goes into the stage of tile cc - 1, which every warp has
    // normally left by now; the warps take turns as feeder
    if (ci.L < ntiles) {
      if (warp == (ci.L % WARPS) && lane == 0) {
        if (ci.L >= STAGES) mbar_wait(&empty[ci.L % STAGES], (uint32_t)((ci.L / STAGES - 1) & 1));
        issue(ci);
      }
      ci.next(plan, walk);
    }
    compute(cc, gc, in);
  };

  // prologue: the first tile alone, so that the first gathers do not queue behind the whole ring's worth of streams
  if (tid == 0) issue(ci);
  ci.next(plan, walk);
  prefetch_state(cg.b);
  Gathered ga, gb;
  gather(cg, ga);
  if (tid == 0) {
    for (int s = 1; s < STAGES - 1 && ci.L < ntiles; ++s) { issue(ci); ci.next(plan, walk); }
  }
  ci.L = min(STAGES - 1, ntiles);                 // every thread's copy of the feeder cursor: re-walk to where lane 0 is
  {
    Cursor t;
    t.init(plan, walk);
    while (t.L < ci.L) t.next(plan, walk);
    ci = t;
  }
  // the two gather buffers alternate by name over a loop unrolled by two (a register rotation would wait for the
  // loads it has just issued)
#pragma unroll 1
  for (;;) {
    step(ga, gb);
    if (cg.L >= ntiles) break;
    step(gb, ga);
    if (cg.L >= ntiles) break;
  }
  gn_flush_segment<float>(plan, ctx.b, ctx.seg_tile, last_tile, acc, partials, red_smem);
}

// ---------------------------------------------------------------------------------------------------------
// generic path (fp64, or fp32 inputs that do not meet the alignment rules of the staged kernel): the same tile
// plan and partial-row layout, plain loads, one correspondence per thread and pass.
// ---------------------------------------------------------------------------------------------------------
template <typename T, bool PT2PL, int DIM, bool FASTW>
__global__ void __launch_bounds__(GN_THREADS)
gn_accum_generic_kernel(const T* __restrict__ source, const T* __restrict__ target, const void* __restrict__ rows,
                        const T* __restrict__ w_init, const int32_t* __restrict__ idx, StateView state, int n, int m, int S,
                        WeightParams P, T* __restrict__ weights_out, long long w_pair_stride, double* __restrict__ partials,
                        TilePlan plan, const int* __restrict__ stop_flag, const int32_t* __restrict__ n_valid) {
  if (stop_flag != nullptr && *stop_flag != 0) return;
  __shared__ double red_smem[(GN_THREADS / 32) * 32];
  constexpr int R = PT2PL ? 1 : 3;
  TileWalk walk;
  walk.init(plan, blockIdx.x);
  const int ntiles = walk.ntiles;
  if (ntiles == 0) return;
  const bool packed = sizeof(T) == 4 && rows != nullptr;
  T acc[RED_N];
  GnPairCtx<T> ctx;
  ctx.b = -1;
  long long last_tile = 0;
  TileCursor c;
  for (c.init(plan, walk); c.L < ntiles; c.next(plan, walk)) {
    const int b = c.b;
    const int i0 = c.j * plan.tile;
    if (c.seg_first) {
      if (c.L > 0) gn_flush_segment<T>(plan, ctx.b, ctx.seg_tile, last_tile, acc, partials, red_smem);
      gn_enter_segment<T>(ctx, b, c.tile(plan), state, acc);
    }
    last_tile = c.tile(plan);
    const T* src_b = source + (size_t)b * n * 3;
    const int32_t* idx_b = idx + (size_t)b * n;
    const int nb = n_valid != nullptr ? min(n_valid[b], n) : n;      // rows >= n_valid[b] are padding: weight 0, never evaluated
    const int i_end = min(i0 + plan.tile, nb);
    if (weights_out != nullptr) {
      for (int i = max(i0, nb) + threadIdx.x; i < min(i0 + plan.tile, n); i += GN_THREADS)
        for (int k = 0; k < R; ++k) weights_out[(size_t)b * w_pair_stride + (size_t)i * R + k] = T(0);
    }
    for (int i = i0 + threadIdx.x; i < i_end; i += GN_THREADS) {
      const int j = idx_b[i];
      const T p[3] = {src_b[3 * i], src_b[3 * i + 1], src_b[3 * i + 2]};
      const T w0 = (w_init != nullptr ? w_init[(size_t)b * n + i] : T(1)) * ctx.active;           // ICP.py:257
      T q[3], nr[3];
      if (packed) load_packed_row<T, PT2PL>(rows, (size_t)b * m + j, q, nr);
      else load_target_row<T, PT2PL>(target + ((size_t)b * m + j) * S, S, q, nr);
      PointEval<T> ev;
      eval_point<T, PT2PL, FASTW>(ctx.C, ctx.r, p, q, nr, w0, P, ev);
      accum_point<T, PT2PL, DIM>(acc, ev, w0, (T)P.thresh);
      if (weights_out != nullptr) {
        T* wrow = weights_out + (size_t)b * w_pair_stride + (size_t)i * R;
#pragma unroll
        for (int k = 0; k < R; ++k) wrow[k] = ev.w;                                               // ICP.py:222, 164-165
      }
    }
  }
  gn_flush_segment<T>(plan, ctx.b, ctx.seg_tile, last_tile, acc, partials, red_smem);
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
    const size_t cap = (size_t)device_sm_count() * 32;
    const int grid = (int)((total + 255) / 256 < cap ? (total + 255) / 256 : cap);
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
                  double* partials, const int* stop_flag, cudaStream_t st, const int32_t* n_valid) {
  const WeightParams P = make_weight_params(p);
  const TilePlan plan = gn_tile_plan(N, n);
  const bool pl = p->icp_type == DICP_PT2PL;
  const bool fast = weight_params_fast(P);
  if (N > 65535) { set_error("gn_accumulate: N > 65535 pairs per call is not supported"); return DICP_ERR_ARG; }
  auto aligned16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  // the staged kernel moves 16-byte aligned slices of the streams with bulk copies
  const bool rows24 = GN_ROWS24 && pl && S == 6 && aligned16(target);      // see load_row24
  const bool unal = (n % 4) != 0;       // the staged kernel then copies aligned supersets of the slices (UNAL)
  const bool staged = sizeof(T) == 4 && (rows24 || rows != nullptr) && aligned16(source) && aligned16(w_init) &&
                      (unal ? !GN_ROWS24 : aligned16(idx));
  if (staged) {
    if constexpr (sizeof(T) == 4) {
      constexpr size_t smem_rest = (GN_THREADS / 32) * (32 + 16) * sizeof(double) + 2 * GN_STAGES * sizeof(uint64_t);
#define DICP_GN_LAUNCH2(PL, DM, FW, RG, UA)                                                                                         \
  do {                                                                                                                             \
    constexpr size_t smem = (size_t)GN_STAGES * (GN_TILE + ((UA) ? 8 : 0)) * 20 + smem_rest;                                       \
    if (smem > 48 * 1024)             /* per device and cheap: no cached flag (a process may drive several devices) */             \
      DICP_CUDA(cudaFuncSetAttribute(gn_accum_staged_kernel<PL, DM, FW, RG, UA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    gn_accum_staged_kernel<PL, DM, FW, RG, UA><<<dim3(plan.ctas, 1, 1), GN_THREADS, smem, st>>>(                                    \
        source, target, (pl && rows24) ? nullptr : rows, w_init, idx, state, n, m, P, weights_out, w_pair_stride, partials, plan,  \
        stop_flag, N, n_valid);                                                                                                    \
  } while (0)
#define DICP_GN_LAUNCH1(PL, DM, FW, RG)            \
  do {                                             \
    if (unal) DICP_GN_LAUNCH2(PL, DM, FW, RG, true); \
    else DICP_GN_LAUNCH2(PL, DM, FW, RG, false);   \
  } while (0)
#define DICP_GN_LAUNCH(PL, DM)                                  \
  do {                                                          \
    if (fast && n_valid == nullptr) DICP_GN_LAUNCH1(PL, DM, true, false); \
    else if (fast) DICP_GN_LAUNCH1(PL, DM, true, true);         \
    else if (n_valid == nullptr) DICP_GN_LAUNCH1(PL, DM, false, false); \
    else DICP_GN_LAUNCH1(PL, DM, false, true);                  \
  } while (0)
      if (pl && p->dim == 3) DICP_GN_LAUNCH(true, 3);
      else if (pl) DICP_GN_LAUNCH(true, 2);
      else if (p->dim == 3) DICP_GN_LAUNCH(false, 3);
      else DICP_GN_LAUNCH(false, 2);
#undef DICP_GN_LAUNCH
#undef DICP_GN_LAUNCH1
#undef DICP_GN_LAUNCH2
    }
  } else {
#define DICP_GN_LAUNCH(PL, DM)                                                                                                     \
  do {                                                                                                                             \
    if (fast) gn_accum_generic_kernel<T, PL, DM, true><<<plan.ctas, GN_THREADS, 0, st>>>(source, target, rows, w_init, idx, state, n, m, S, P, \
                                                                                       weights_out, w_pair_stride, partials, plan, stop_flag, n_valid); \
    else gn_accum_generic_kernel<T, PL, DM, false><<<plan.ctas, GN_THREADS, 0, st>>>(source, target, rows, w_init, idx, state, n, m, S, P, \
                                                                                   weights_out, w_pair_stride, partials, plan, stop_flag, n_valid); \
  } while (0)
    if (pl && p->dim == 3) DICP_GN_LAUNCH(true, 3);
    else if (pl) DICP_GN_LAUNCH(true, 2);
    else if (p->dim == 3) DICP_GN_LAUNCH(false, 3);
    else DICP_GN_LAUNCH(false, 2);
#undef DICP_GN_LAUNCH
  }
  DICP_LAUNCH_CHECK();
  return 0;
}
template int gn_accumulate<float>(const float*, const float*, const void*, const float*, const int32_t*, StateView, int, int, int, int,
                                  const dicp_params*, float*, long long, double*, const int*, cudaStream_t, const int32_t*);
template int gn_accumulate<double>(const double*, const double*, const void*, const double*, const int32_t*, StateView, int, int, int, int,
                                   const dicp_params*, double*, long long, double*, const int*, cudaStream_t, const int32_t*);

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
                T* __restrict__ num_iters, T* __restrict__ match_ratio, const int32_t* __restrict__ n_valid,
                double* status_rec) {
  const int b = blockIdx.y;
  const int nb = n_valid != nullptr ? min(n_valid[b], n) : n;      // padding rows of a ragged batch are zero points: pc = r
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
    T a[3] = {T(0), T(0), T(0)};
    if (i < nb) {
      const T p[3] = {src_b[3 * i], src_b[3 * i + 1], src_b[3 * i + 2]};
      rotate(C, p, a);
    }
    pc_b[3 * i] = a[0] + r[0]; pc_b[3 * i + 1] = a[1] + r[1]; pc_b[3 * i + 2] = a[2] + r[2];   // ICP.py:274
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && b == 0) {      // the call's status, for callers that did not wait for it
    status_rec[DICP_STATE_ITERS_RUN] = (double)kr;
    status_rec[DICP_STATE_SINGULAR] = ctrl->singular ? 1.0 : 0.0;
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
IcpWorkspace plan_workspace(void* ws, size_t ws_bytes, int N, int n, int m, int elem_size, bool backward_only) {
  Arena ar(ws, ws_bytes);
  IcpWorkspace w;
  // the reverse pass needs none of the search structures: its workspace is a fraction of the forward's, so the caller
  // can drop the forward workspace as soon as the forward call returns
  w.packed = ar.take<char>(backward_only ? 1 : nn_packed_bytes(N, m, elem_size));
  w.nn_ws_bytes = backward_only ? 0 : nn_search_ws_bytes(N, n, m, elem_size);
  w.nn_ws = ar.take<char>(w.nn_ws_bytes > 0 ? w.nn_ws_bytes : 1);
  w.partials = ar.take<double>(backward_only ? 1 : (size_t)N * gn_blocks_per_pair(N, n) * RED_N);
  w.ctrl = ar.take<Ctrl>(1);
  w.bw_state = ar.take<double>(backward_only ? (size_t)N * BW_STRIDE : 1);
  w.bw_partials = ar.take<double>(backward_only ? (size_t)N * bw_rows_per_pair(N, n) * BW_NRED : 1);
  w.grid_ws_bytes = backward_only ? 0 : grid_ws_layout(nullptr, 0, N, n, m, elem_size).total;
  w.grid_ws = ar.take<char>(w.grid_ws_bytes > 0 ? w.grid_ws_bytes : 1);
  w.rows = ar.take<char>(gn_rows_bytes(N, m, elem_size) > 0 ? gn_rows_bytes(N, m, elem_size) : 1);
  w.gtab = ar.take<char>(backward_only && bw_gtab_bytes(N, m, elem_size) > 0 ? bw_gtab_bytes(N, m, elem_size) : 1);
  w.total = ar.off;
  return w;
}

template <typename T>
static int icp_forward_impl(const T* source, const T* target, const T* T_init, const T* w_init, const int32_t* m_valid,
                            const int32_t* n_valid, int N, int n,
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
  if (ws == nullptr || ws_bytes < plan_workspace(nullptr, 0, N, n, m, sizeof(T), false).total) {
    set_error("dicp_icp_forward: workspace too small");
    return DICP_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  typedef typename Vec4<T>::type V4;
  IcpWorkspace w = plan_workspace(ws, ws_bytes, N, n, m, sizeof(T), false);
  const int K = p->max_iterations;
  const bool pl = p->icp_type == DICP_PT2PL;
  const long long nw = (long long)n * (pl ? 1 : 3);
  const int nblk = gn_blocks_per_pair(N, n);
  const int check_every = p->check_every >= 1 ? p->check_every : 1;
  const size_t rec_stride = (size_t)N * ST_STRIDE;
  // DICP_FLAG_NO_SYNC: a constant-iteration loop needs nothing from the device before it returns
  const bool nosync = p->const_iter && (p->flags & DICP_FLAG_NO_SYNC);

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
    if (!nosync) {
      int max_occ = 0;
      DICP_CUDA(cudaMemcpyAsync(&max_occ, gw.max_occupancy, sizeof(int), cudaMemcpyDeviceToHost, st));
      DICP_CUDA(cudaStreamSynchronize(st));
      if (max_occ > 64 && max_occ > m / 64) use_grid = false;
    }
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
    } else if (use_grid) {      // GRID_WARM (off): from the second iteration on the previous neighbours would bound the search
      rc = grid_search<T>(source, rec_k, N, n, m, gw, true, idx_k, nullptr, &w.ctrl->stop, st, target, S,
                          (GRID_WARM && k > 0) ? idx_k - (size_t)N * n : nullptr);
    }
    else rc = nn_search<T>(source, rec_k, (const V4*)w.packed, N, n, m, idx_k, nullptr, &w.ctrl->stop, w.nn_ws, w.nn_ws_bytes, st, n_valid);
    if (rc) return rc;
    rc = gn_accumulate<T>(source, target, rows, w_init, idx_k, state_view_records(rec_k), N, n, m, S, p,
                          weights != nullptr ? weights + (size_t)k * nw : nullptr, (long long)K * nw, w.partials, &w.ctrl->stop, st, n_valid);
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
  dim3 fgrid((n + 2047) / 2048 < 128 ? (n + 2047) / 2048 : 128, N);
  finalize_kernel<T><<<fgrid, 256, 0, st>>>(source, n, saved_state, N, w.ctrl, p->match_ratio_thresh, (double)nw,
                                            pc, T_out, num_iters, match_ratio, n_valid, saved_state);
  DICP_LAUNCH_CHECK();
  if (nosync) {                                   // the status stays in saved_state[0][0][DICP_STATE_ITERS_RUN / _SINGULAR]
    *iters_run_host = K;
    return 0;
  }
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
  return dicp::plan_workspace(nullptr, 0, N, n, m, elem_size, false).total;
}
size_t dicp_icp_backward_workspace_bytes(int N, int n, int m, int elem_size) {
  return dicp::plan_workspace(nullptr, 0, N, n, m, elem_size, true).total;
}

int dicp_icp_forward_f32(const float* source, const float* target, const float* T_init, const float* w_init, const int32_t* m_valid, const int32_t* n_valid,
                         int N, int n, int m, int S, const dicp_params* params, float* pc, float* T_out, float* deltas, float* weights, float* costs,
                         uint8_t* converged, float* num_iters, float* match_ratio, int32_t* saved_idx, double* saved_state,
                         int32_t* iters_run_host, void* ws, size_t ws_bytes, void* stream) {
  return dicp::icp_forward_impl<float>(source, target, T_init, w_init, m_valid, n_valid, N, n, m, S, params, pc, T_out, deltas, weights, costs,
                                       converged, num_iters, match_ratio, saved_idx, saved_state, iters_run_host, ws, ws_bytes, stream);
}
int dicp_icp_forward_f64(const double* source, const double* target, const double* T_init, const double* w_init, const int32_t* m_valid, const int32_t* n_valid,
                         int N, int n, int m, int S, const dicp_params* params, double* pc, double* T_out, double* deltas, double* weights,
                         double* costs, uint8_t* converged, double* num_iters, double* match_ratio, int32_t* saved_idx,
                         double* saved_state, int32_t* iters_run_host, void* ws, size_t ws_bytes, void* stream) {
  return dicp::icp_forward_impl<double>(source, target, T_init, w_init, m_valid, n_valid, N, n, m, S, params, pc, T_out, deltas, weights, costs,
                                        converged, num_iters, match_ratio, saved_idx, saved_state, iters_run_host, ws, ws_bytes, stream);
}
int dicp_icp_forward_hooked_f32(const float* source, const float* target, const float* T_init, const float* w_init, const int32_t* m_valid, const int32_t* n_valid,
                                int N, int n, int m, int S, const dicp_params* params, float* pc, float* T_out, float* deltas, float* weights,
                                float* costs, uint8_t* converged, float* num_iters, float* match_ratio, int32_t* saved_idx,
                                double* saved_state, int32_t* iters_run_host, void* ws, size_t ws_bytes, void* stream, dicp_nn_hook hook,
                                void* user) {
  if (hook == nullptr) { dicp::set_error("dicp_icp_forward_hooked: null hook"); return DICP_ERR_ARG; }
  return dicp::icp_forward_impl<float>(source, target, T_init, w_init, m_valid, n_valid, N, n, m, S, params, pc, T_out, deltas, weights, costs,
                                       converged, num_iters, match_ratio, saved_idx, saved_state, iters_run_host, ws, ws_bytes, stream, hook, user);
}
int dicp_icp_forward_hooked_f64(const double* source, const double* target, const double* T_init, const double* w_init, const int32_t* m_valid, const int32_t* n_valid,
                                int N, int n, int m, int S, const dicp_params* params, double* pc, double* T_out, double* deltas, double* weights,
                                double* costs, uint8_t* converged, double* num_iters, double* match_ratio, int32_t* saved_idx,
                                double* saved_state, int32_t* iters_run_host, void* ws, size_t ws_bytes, void* stream, dicp_nn_hook hook,
                                void* user) {
  if (hook == nullptr) { dicp::set_error("dicp_icp_forward_hooked: null hook"); return DICP_ERR_ARG; }
  return dicp::icp_forward_impl<double>(source, target, T_init, w_init, m_valid, n_valid, N, n, m, S, params, pc, T_out, deltas, weights, costs,
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
int dicp_gn_blocks_per_pair(int N, int n) { return N >= 1 && n >= 1 ? dicp::gn_blocks_per_pair(N, n) : 0; }
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
