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
// HBM-bound: 124 B per correspondence (pt2pl).  The point pass runs on the persistent tile schedule and the
// TMA-fed staging ring of tile_stream.cuh (see gn_kernels.cu for the forward twin); its scatter goes to a
// 16-byte aligned gradient table (8 floats per target row, 4 for point-to-point) with 128-bit vector reductions
// (red.global.add.v4.f32: two per correspondence instead of three 64-bit ones on rows that straddle sectors), which
// one pass folds into the caller's [N,m,S] gradient at the end of the reverse loop.
#include "common.cuh"
#include "tile_stream.cuh"

namespace dicp {

constexpr int BW_THREADS = 128;
#ifndef BW_U
#define BW_U 2
#endif
#ifndef BW_STAGES
#define BW_STAGES 4
#endif
#ifndef BW_MINB
#define BW_MINB 3
#endif
#ifndef BW_K
#define BW_K 2            // 512-correspondence chunks: the read-modify-write of the gradient table wants tight locality (k = 8: 641 us, k = 2: 461 us at 1024 pairs)
#endif
#ifndef BW_GTAB
#define BW_GTAB 1          // 1: scatter into the aligned gradient table (two red.v4.f32), 0: into the caller's rows (three red.v2.f32)
#endif
#ifndef BW_PDL
#define BW_PDL 2           // programmatic dependent launches in the reverse loop.  bit 1: the point pass starts while the pair kernel runs
                           // (its prologue -- barriers, first streams, first row gathers -- needs nothing from it): 141.9 -> 137.4 us per
                           // reverse iteration at C3.  bit 0: the pair kernel as a dependent of the point pass: 156 us (worse), off
#endif
constexpr int BW_TILE = BW_THREADS * BW_U;
constexpr int BW_STAGE_BYTES = BW_TILE * 36;      // idx 4 + p 12 + w0 4 + grad_source 12 + grad_w 4 per correspondence

TilePlan bw_tile_plan(int N, int n) { return make_tile_plan(N, n, BW_TILE, BW_MINB, BW_K); }
int bw_seed_blocks(int n) {                      // the seed pass: one partial row per CTA, up to 128 CTAs per pair
  int nb = (n + 1023) / 1024;
  return nb < 1 ? 1 : (nb > 128 ? 128 : nb);
}
int bw_rows_per_pair(int N, int n) {             // rows of the [N, rows, BW_NRED] partial-sum table
  const int a = bw_tile_plan(N, n).rows, b = bw_seed_blocks(n);
  return a > b ? a : b;
}
size_t bw_gtab_bytes(int N, int m, int elem_size) { return elem_size == 4 ? (size_t)N * m * 32 : 0; }

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
                T* __restrict__ grad_source, double* __restrict__ partials, const int32_t* __restrict__ n_valid) {
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
    // padding rows of a ragged batch (i >= n_valid[b]) are zero points whatever the buffer holds
    const bool real = n_valid == nullptr || i < n_valid[b];
    const T p[3] = {real ? src_b[3 * i] : T(0), real ? src_b[3 * i + 1] : T(0), real ? src_b[3 * i + 2] : T(0)};
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
// One warp per pair: lanes 0..11 each own one component of the carry (Cbar 9, rbar 3) and add up the partial rows
// (coalesced 96-byte row reads, fixed order), lane 0 then runs the serial 3x3 / 6x6 algebra in double.
template <typename T>
__global__ void __launch_bounds__(128)
bwd_pair_kernel(const double* __restrict__ partials, int nblk, int row_stride, const double* __restrict__ rec_k, int dim, int mode,
                const T* __restrict__ grad_T, double* __restrict__ bw_state, T* __restrict__ grad_T_init, int N) {
  // PDL: this kernel may have been launched while the point pass whose rows it sums was still running, and the next
  // point pass may start its prologue (streams, first row gathers) while this kernel runs -- it waits before it reads
  // bw_state.  The trigger comes after the wait, so whatever follows us also sees that point pass completed.
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= N) return;
  double* bw = bw_state + (size_t)b * BW_STRIDE;
  double mine = 0.0;
  if (lane < 12) {
    if (mode == 0) {
      const int rr = lane < 9 ? lane / 3 : lane - 9, c = lane < 9 ? lane % 3 : 3;
      mine = grad_T != nullptr ? (double)grad_T[(size_t)b * 16 + rr * 4 + c] : 0.0;
    } else {
      mine = bw[lane];
    }
  }
  {
    // rows summed by two groups of 12 lanes (even / odd rows), four independent loads in flight per lane: fixed order
    const int comp = lane % 12, grp = lane / 12;
    double part = 0.0;
    if (lane < 24) {
      const double* rows_b = partials + (size_t)b * row_stride * BW_NRED + comp;
      int j = grp;
      for (; j + 6 < nblk; j += 8) {
        const double a0 = rows_b[(size_t)j * BW_NRED], a1 = rows_b[(size_t)(j + 2) * BW_NRED], a2 = rows_b[(size_t)(j + 4) * BW_NRED],
                     a3 = rows_b[(size_t)(j + 6) * BW_NRED];
        part += a0; part += a1; part += a2; part += a3;
      }
      for (; j < nblk; j += 2) part += rows_b[(size_t)j * BW_NRED];
    }
    const double other = __shfl_down_sync(0xffffffffu, part, 12);
    if (lane < 12) mine += part + other;
  }
  if (mode == 2) {
    if (lane < 12) {
      const int rr = lane < 9 ? lane / 3 : lane - 9, c = lane < 9 ? lane % 3 : 3;
      grad_T_init[(size_t)b * 16 + rr * 4 + c] = (T)mine;
    } else if (lane < 16) {
      grad_T_init[(size_t)b * 16 + 12 + (lane - 12)] = T(0);
    }
    return;
  }
  double carry[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) carry[i] = __shfl_sync(0xffffffffu, mine, i);
  if (lane != 0) return;
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

// per-segment hand-over of the 12 partial sums (sum abar p^T, sum ebar), same row numbering as the forward
template <typename T>
__device__ __forceinline__ void bw_flush_segment(const TilePlan& plan, int nrows, int b, long long seg_tile, long long last_tile,
                                                 const T* acc, double* __restrict__ partials, double* red_smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int row = tp_segment_row(plan, b, seg_tile);
  double* rows_b = partials + (size_t)b * nrows * BW_NRED;
  T v[32];
#pragma unroll
  for (int i = 0; i < BW_NRED; ++i) v[i] = acc[i];
  warp_multi_reduce<T, BW_NRED>(v);
  red_smem[warp * 32 + lane] = (double)v[0];
  __syncthreads();
  if (threadIdx.x < BW_NRED) {
    double s = 0.0;
    for (int w = 0; w < nwarps; ++w) s += red_smem[w * 32 + threadIdx.x];
    rows_b[(size_t)row * BW_NRED + threadIdx.x] = s;
  }
  if (last_tile == (long long)b * plan.tpp + plan.tpp - 1) {
    for (int i = (row + 1) * BW_NRED + threadIdx.x; i < nrows * BW_NRED; i += blockDim.x) rows_b[i] = 0.0;
  }
  __syncthreads();
}

// what the point pass needs to know about a pair in iteration k
template <typename T>
struct BwPairCtx {
  T C[9], r[3], lam[6], del[6];
  bool active;
  int b;
  long long seg_tile;
};

// scatter-add of one correspondence's target gradient into the aligned table (fp32 fast path)
template <bool PT2PL>
__device__ __forceinline__ void scatter_gtab(float* __restrict__ gtab, size_t row, const float* ebar, const float* nbar) {
  if (PT2PL) {
    float4* g = reinterpret_cast<float4*>(gtab) + row * 2;
    atomicAdd(g, make_float4(-ebar[0], -ebar[1], -ebar[2], nbar[0]));
    atomicAdd(g + 1, make_float4(nbar[1], nbar[2], 0.0f, 0.0f));
  } else {
    atomicAdd(reinterpret_cast<float4*>(gtab) + row, make_float4(-ebar[0], -ebar[1], -ebar[2], 0.0f));
  }
}

// ---------------------------------------------------------------------------------------------------------
// fp32 fast path (16-byte aligned stream bases, gather table `rows` and gradient table `gtab` present; UNAL when n % 4 != 0).
// Ring stage = { idx | p | w0 | grad_source | grad_w } slices of the tile; see gn_accum_staged_kernel for the
// full / empty barrier protocol, the rotating feeder and the state prefetch.
// ---------------------------------------------------------------------------------------------------------
template <bool PT2PL, bool FASTW, bool RAGGED, bool UNAL>
__global__ void __launch_bounds__(BW_THREADS, BW_MINB)
bwd_points_staged_kernel(const float* __restrict__ source, const void* __restrict__ rows, const float* __restrict__ w_init,
                         const int32_t* __restrict__ idx, const double* __restrict__ rec_k, const double* __restrict__ bw_state,
                         int n, int m, WeightParams P, float* __restrict__ grad_source, float* __restrict__ gtab,
                         float* __restrict__ grad_w, double* __restrict__ partials, int nrows, TilePlan plan, int gS,
                         const int32_t* __restrict__ n_valid) {
  constexpr int U = BW_U, TILE = BW_TILE, STAGES = BW_STAGES, WARPS = BW_THREADS / 32;
  static_assert(STAGES >= 3, "a stage is refilled one step after its release");
  extern __shared__ __align__(128) unsigned char bw_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // UNAL: clouds whose length is not a multiple of 4 -- the feeder copies the aligned superset of a tile's slice of every
  // stream and the consumers read at the slice's offset inside it (see gn_accum_staged_kernel)
  constexpr int STG = TILE + (UNAL ? 8 : 0);
  constexpr size_t STAGE_BYTES = (size_t)STG * 36;
  auto st_idx = [&](int s) { return reinterpret_cast<int32_t*>(bw_smem + (size_t)s * STAGE_BYTES); };
  auto st_p = [&](int s) { return reinterpret_cast<float*>(bw_smem + (size_t)s * STAGE_BYTES + STG * 4); };
  auto st_w = [&](int s) { return reinterpret_cast<float*>(bw_smem + (size_t)s * STAGE_BYTES + STG * 16); };
  auto st_gs = [&](int s) { return reinterpret_cast<float*>(bw_smem + (size_t)s * STAGE_BYTES + STG * 20); };
  auto st_gw = [&](int s) { return reinterpret_cast<float*>(bw_smem + (size_t)s * STAGE_BYTES + STG * 32); };
  auto st_off = [&](const TileCursor& c) { return UNAL ? ((c.b & 3) * (n & 3)) & 3 : 0; };      // (b n + j TILE) mod 4
  const int idx_mis = UNAL ? (int)((reinterpret_cast<uintptr_t>(idx) >> 2) & 3) : 0;              // elements past a 16-byte boundary
  double* red_smem = reinterpret_cast<double*>(bw_smem + (size_t)STAGES * STAGE_BYTES);
  double* slot = red_smem + WARPS * 32 + warp * 32;            // this warp's copy of the next segment's pair record
  uint64_t* full = reinterpret_cast<uint64_t*>(red_smem + WARPS * 32 + WARPS * 32);
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

  typedef TileCursor Cursor;
  auto tile_count = [&](const Cursor& c) { return min(TILE, n - c.j * TILE); };
  // real correspondences of the tile: rows >= n_valid[b] of a ragged batch are padding, nothing flows into them
  // (a compile-time option, as in gn_accum_staged_kernel)
  auto tile_valid = [&](const Cursor& c) {
    if (!RAGGED) return min(TILE, n - c.j * TILE);
    return max(0, min(TILE, min(n_valid[c.b], n) - c.j * TILE));
  };
  auto issue = [&](const Cursor& c) {
    const int s = c.L % STAGES;
    const uint32_t cnt = (uint32_t)tile_count(c);
    const size_t o = (size_t)c.b * n + c.j * TILE;
    if (UNAL) {
      // aligned supersets of the slices, tails behind the last whole 16-byte group by ordinary copies: see gn_accum_staged_kernel
      const long long total = (long long)(plan.tiles / plan.tpp) * n, oo = (long long)o;
      auto span = [&](int mis, long long& a0, uint32_t& nb) {
        a0 = ((oo + mis) & ~3LL) - mis;
        long long a1 = ((oo + cnt + mis + 3) & ~3LL) - mis;
        const long long lim = ((total + mis) & ~3LL) - mis;
        if (a1 > lim) a1 = lim;
        nb = a1 > a0 ? (uint32_t)(a1 - a0) : 0u;
      };
      long long a0, ai0;
      uint32_t nb, nbi;
      span(0, a0, nb);
      span(idx_mis, ai0, nbi);
      for (long long e = (a0 + nb > oo ? a0 + nb : oo); e < oo + cnt; ++e) {
        const long long k = e - a0;
#pragma unroll
        for (int c3 = 0; c3 < 3; ++c3) { st_p(s)[k * 3 + c3] = source[e * 3 + c3]; st_gs(s)[k * 3 + c3] = grad_source[e * 3 + c3]; }
        if (w_init != nullptr) st_w(s)[k] = w_init[e];
        if (grad_w != nullptr) st_gw(s)[k] = grad_w[e];
      }
      for (long long e = (ai0 + nbi > oo ? ai0 + nbi : oo); e < oo + cnt; ++e) st_idx(s)[e - ai0] = idx[e];
      mbar_expect_tx(&full[s], nbi * 4u + nb * (24u + (w_init != nullptr ? 4u : 0u) + (grad_w != nullptr ? 4u : 0u)));
      if (nbi > 0u) bulk_g2s(st_idx(s), idx + ai0, nbi * 4u, &full[s]);
      if (nb > 0u) {
        bulk_g2s(st_p(s), source + a0 * 3, nb * 12u, &full[s]);
        bulk_g2s(st_gs(s), grad_source + a0 * 3, nb * 12u, &full[s]);
        if (w_init != nullptr) bulk_g2s(st_w(s), w_init + a0, nb * 4u, &full[s]);
        if (grad_w != nullptr) bulk_g2s(st_gw(s), grad_w + a0, nb * 4u, &full[s]);
      }
      return;
    }
    mbar_expect_tx(&full[s], cnt * (28u + (w_init != nullptr ? 4u : 0u) + (grad_w != nullptr ? 4u : 0u)));
    bulk_g2s(st_idx(s), idx + o, cnt * 4u, &full[s]);
    bulk_g2s(st_p(s), source + o * 3, cnt * 12u, &full[s]);
    bulk_g2s(st_gs(s), grad_source + o * 3, cnt * 12u, &full[s]);
    if (w_init != nullptr) bulk_g2s(st_w(s), w_init + o, cnt * 4u, &full[s]);
    if (grad_w != nullptr) bulk_g2s(st_gw(s), grad_w + o, cnt * 4u, &full[s]);
  };
  // (C 9, r 3) of the iteration and (lambda 6, delta 6, active) of the reverse step into the warp's slot
  auto prefetch_state = [&](int b) {
    const double* rec = rec_k + (size_t)b * ST_STRIDE;
    const double* bw = bw_state + (size_t)b * BW_STRIDE;
    if (lane < 12) cp_async8(slot + lane, rec + (lane < 9 ? ST_C + lane : ST_R + lane - 9));
    else if (lane < 24) cp_async8(slot + lane, bw + BW_LAM + (lane - 12));       // BW_LAM .. BW_DEL + 5 are contiguous
    else if (lane == 24) cp_async8(slot + 24, bw + BW_ACTIVE);
  };
  static_assert(BW_DEL == BW_LAM + 6, "lambda and delta are fetched as one run");

  struct Gathered { float q[U][3], nr[U][3]; int jj[U]; };
  auto gather = [&](const Cursor& c, Gathered& g) {
    const int s = c.L % STAGES;
    const int cnt = tile_valid(c);
    mbar_wait(&full[s], (uint32_t)((c.L / STAGES) & 1));
    const size_t row_b = (size_t)c.b * m;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = tid * U + u < cnt ? st_idx(s)[((st_off(c) + idx_mis) & 3) + tid * U + u] : 0;
      g.jj[u] = j;
      load_packed_row<float, PT2PL>(rows, row_b + j, g.q[u], g.nr[u]);
    }
  };

  float acc[BW_NRED];
  BwPairCtx<float> ctx;
  ctx.b = -1;
  ctx.active = false;
  long long last_tile = 0;
  auto enter = [&](const Cursor& c) {
    if (c.L > 0) bw_flush_segment<float>(plan, nrows, ctx.b, ctx.seg_tile, last_tile, acc, partials, red_smem);
    cp_async_wait_all();
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 9; ++i) ctx.C[i] = (float)slot[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) ctx.r[i] = (float)slot[9 + i];
#pragma unroll
    for (int i = 0; i < 6; ++i) { ctx.lam[i] = (float)slot[12 + i]; ctx.del[i] = (float)slot[18 + i]; }
    ctx.active = slot[24] != 0.0;
    ctx.b = c.b;
    ctx.seg_tile = c.tile(plan);
#pragma unroll
    for (int i = 0; i < BW_NRED; ++i) acc[i] = 0.0f;
    __syncwarp();
  };

  Cursor ci, cg;
  ci.init(plan, walk);
  cg.init(plan, walk);
  auto step = [&](const Gathered& gc, Gathered& gn) {
    const Cursor cc = cg;
    if (cc.seg_first) enter(cc);
    last_tile = cc.tile(plan);
    const int s = cc.L % STAGES;
    const int cnt = tile_valid(cc);
    // this lane's staged inputs
    float p[U][3], gs[U][3], w0[U], gw[U];
    const int off = st_off(cc);
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int c = 0; c < 3; ++c) { p[u][c] = st_p(s)[(off + tid * U + u) * 3 + c]; gs[u][c] = st_gs(s)[(off + tid * U + u) * 3 + c]; }
      w0[u] = w_init != nullptr ? st_w(s)[off + tid * U + u] : 1.0f;
      gw[u] = grad_w != nullptr ? st_gw(s)[off + tid * U + u] : 0.0f;
    }
    cg.next(plan, walk);
    if (cg.L < ntiles) {
      if (cg.seg_first) prefetch_state(cg.b);
      gather(cg, gn);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
    if (ci.L < ntiles) {
      if (warp == (ci.L % WARPS) && lane == 0) {
        if (ci.L >= STAGES) mbar_wait(&empty[ci.L % STAGES], (uint32_t)((ci.L / STAGES - 1) & 1));
        issue(ci);
      }
      ci.next(plan, walk);
    }
    if (!ctx.active) return;                  // frozen pair: nothing flows into its points (ICP.py:256-257)
    const size_t o = (size_t)cc.b * n + cc.j * TILE + tid * U;
    const size_t row_b = (size_t)cc.b * m;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (tid * U + u >= cnt) continue;
      PointEval<float> ev;
      eval_point<float, PT2PL, FASTW>(ctx.C, ctx.r, p[u], gc.q[u], gc.nr[u], w0[u], P, ev);
      float abar[3], ebar[3], nbar[3], w0bar;
      point_backward<float, PT2PL>(ev, w0[u], ctx.lam, ctx.del, abar, ebar, nbar, w0bar);
      if (grad_w != nullptr) grad_w[o + u] = gw[u] + w0bar;
#pragma unroll
      for (int c = 0; c < 3; ++c)
        grad_source[(o + u) * 3 + c] = gs[u][c] + (ctx.C[c] * abar[0] + ctx.C[3 + c] * abar[1] + ctx.C[6 + c] * abar[2]);   // C^T abar
      if (BW_GTAB) {
        scatter_gtab<PT2PL>(gtab, row_b + gc.jj[u], ebar, nbar);
      } else {                                 // gtab is the caller's [N,m,S] gradient here, S = 6 | 3
        float* grow = gtab + (row_b + gc.jj[u]) * (PT2PL ? 6 : gS);
        if (PT2PL) {
          const float g6[6] = {-ebar[0], -ebar[1], -ebar[2], nbar[0], nbar[1], nbar[2]};
          atomic_add_row6(grow, g6);
        } else {
#pragma unroll
          for (int c = 0; c < 3; ++c) atomicAdd(grow + c, -ebar[c]);
        }
      }
#pragma unroll
      for (int rr = 0; rr < 3; ++rr) {
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[rr * 3 + c] += abar[rr] * p[u][c];
        acc[9 + rr] += ebar[rr];
      }
    }
  };

  // PDL: everything up to here and the prologue below read only what kernels before the preceding bwd_pair_kernel
  // wrote (saved correspondences, clouds, gather table, the running gradients of the previous point pass, which that
  // pair kernel waited for before it let this grid start); the pair's (lambda, delta, carry) are read after pdl_wait()
  pdl_trigger();
  if (tid == 0) issue(ci);
  ci.next(plan, walk);
  Gathered ga, gb;
  gather(cg, ga);
  if (tid == 0) {
    for (int s = 1; s < STAGES - 1 && ci.L < ntiles; ++s) { issue(ci); ci.next(plan, walk); }
  }
  pdl_wait();
  prefetch_state(cg.b);
  {
    const int target_L = min(STAGES - 1, ntiles);
    Cursor t;
    t.init(plan, walk);
    while (t.L < target_L) t.next(plan, walk);
    ci = t;
  }
#pragma unroll 1
  for (;;) {
    step(ga, gb);
    if (cg.L >= ntiles) break;
    step(gb, ga);
    if (cg.L >= ntiles) break;
  }
  bw_flush_segment<float>(plan, nrows, ctx.b, ctx.seg_tile, last_tile, acc, partials, red_smem);
}

// ---------------------------------------------------------------------------------------------------------
// generic path (fp64, or fp32 inputs that do not meet the alignment rules): same schedule and partial rows,
// plain loads, scatter straight into the caller's gradient rows.
// ---------------------------------------------------------------------------------------------------------
template <typename T, bool PT2PL, bool FASTW>
__global__ void __launch_bounds__(BW_THREADS)
bwd_points_generic_kernel(const T* __restrict__ source, const T* __restrict__ target, const T* __restrict__ w_init,
                          const int32_t* __restrict__ idx, const double* __restrict__ rec_k, const double* __restrict__ bw_state,
                          int n, int m, int S, WeightParams P, T* __restrict__ grad_source, T* __restrict__ grad_target,
                          T* __restrict__ grad_w, double* __restrict__ partials, int nrows, TilePlan plan,
                          const int32_t* __restrict__ n_valid) {
  __shared__ double red_smem[(BW_THREADS / 32) * 32];
  TileWalk walk;
  walk.init(plan, blockIdx.x);
  const int ntiles = walk.ntiles;
  if (ntiles == 0) return;
  T acc[BW_NRED];
  BwPairCtx<T> ctx;
  ctx.b = -1;
  ctx.active = false;
  long long last_tile = 0;
  TileCursor c;
  for (c.init(plan, walk); c.L < ntiles; c.next(plan, walk)) {
    const int b = c.b;
    if (c.seg_first) {
      if (c.L > 0) bw_flush_segment<T>(plan, nrows, ctx.b, ctx.seg_tile, last_tile, acc, partials, red_smem);
      const double* rec = rec_k + (size_t)b * ST_STRIDE;
      const double* bw = bw_state + (size_t)b * BW_STRIDE;
#pragma unroll
      for (int i = 0; i < 9; ++i) ctx.C[i] = (T)rec[ST_C + i];
#pragma unroll
      for (int i = 0; i < 3; ++i) ctx.r[i] = (T)rec[ST_R + i];
#pragma unroll
      for (int i = 0; i < 6; ++i) { ctx.lam[i] = (T)bw[BW_LAM + i]; ctx.del[i] = (T)bw[BW_DEL + i]; }
      ctx.active = bw[BW_ACTIVE] != 0.0;
      ctx.b = b;
      ctx.seg_tile = c.tile(plan);
#pragma unroll
      for (int i = 0; i < BW_NRED; ++i) acc[i] = T(0);
    }
    last_tile = c.tile(plan);
    if (!ctx.active) continue;
    const T* src_b = source + (size_t)b * n * 3;
    const T* tgt_b = target + (size_t)b * m * S;
    T* gt_b = grad_target + (size_t)b * m * S;
    T* gs_b = grad_source + (size_t)b * n * 3;
    const int i0 = c.j * plan.tile, i_end = min(i0 + plan.tile, n_valid != nullptr ? min(n_valid[b], n) : n);
    for (int i = i0 + threadIdx.x; i < i_end; i += BW_THREADS) {
      const int j = idx[(size_t)b * n + i];
      const T p[3] = {src_b[3 * i], src_b[3 * i + 1], src_b[3 * i + 2]};
      const T w0 = w_init != nullptr ? w_init[(size_t)b * n + i] : T(1);
      T q[3], nr[3];
      load_target_row<T, PT2PL>(tgt_b + (size_t)j * S, S, q, nr);
      PointEval<T> ev;
      eval_point<T, PT2PL, FASTW>(ctx.C, ctx.r, p, q, nr, w0, P, ev);
      T abar[3], ebar[3], nbar[3], w0bar;
      point_backward<T, PT2PL>(ev, w0, ctx.lam, ctx.del, abar, ebar, nbar, w0bar);
      if (grad_w != nullptr) grad_w[(size_t)b * n + i] += w0bar;
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) gs_b[3 * i + cc] += ctx.C[cc] * abar[0] + ctx.C[3 + cc] * abar[1] + ctx.C[6 + cc] * abar[2];   // C^T abar
      T* grow = gt_b + (size_t)j * S;
      if (PT2PL) {
        const T g6[6] = {-ebar[0], -ebar[1], -ebar[2], nbar[0], nbar[1], nbar[2]};
        atomic_add_row6(grow, g6);
      } else {
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) atomicAdd(grow + cc, -ebar[cc]);
      }
#pragma unroll
      for (int rr = 0; rr < 3; ++rr) {
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) acc[rr * 3 + cc] += abar[rr] * p[cc];
        acc[9 + rr] += ebar[rr];
      }
    }
  }
  bw_flush_segment<T>(plan, nrows, ctx.b, ctx.seg_tile, last_tile, acc, partials, red_smem);
}

// folds the aligned gradient table into the caller's [N,m,S] rows (every entry is written)
template <bool PT2PL>
__global__ void __launch_bounds__(256)
bwd_unpack_gtab_kernel(const float* __restrict__ gtab, size_t total, int S, float* __restrict__ grad_target) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    float* g = grad_target + i * S;
    if (PT2PL) {
      const float4 a = reinterpret_cast<const float4*>(gtab)[2 * i], b = reinterpret_cast<const float4*>(gtab)[2 * i + 1];
      *reinterpret_cast<float2*>(g) = make_float2(a.x, a.y);
      *reinterpret_cast<float2*>(g + 2) = make_float2(a.z, a.w);
      *reinterpret_cast<float2*>(g + 4) = make_float2(b.x, b.y);
    } else {
      const float4 a = reinterpret_cast<const float4*>(gtab)[i];
      g[0] = a.x; g[1] = a.y; g[2] = a.z;
      for (int c = 3; c < S; ++c) g[c] = 0.0f;
    }
  }
}

template <typename T>
static int icp_backward_impl(const T* source, const T* target, const T* w_init, const int32_t* n_valid, int N, int n, int m, int S, const dicp_params* p,
                             int iters_run, const int32_t* saved_idx, const double* saved_state, const T* grad_T, const T* grad_pc,
                             T* grad_source, T* grad_target, T* grad_w, T* grad_T_init, void* ws, size_t ws_bytes, void* stream) {
  DICP_CHECK_ARG(source && target && p && saved_idx && saved_state && grad_source && grad_target && grad_T_init, "null pointer");
  DICP_CHECK_ARG(N >= 1 && N <= 65535 && n >= 1 && m >= 1 && (S == 3 || S == 6), "bad sizes");
  DICP_CHECK_ARG(p->icp_type != DICP_PT2PL || S == 6, "pt2pl needs a 6-column target");
  DICP_CHECK_ARG(iters_run >= 1 && iters_run <= p->max_iterations, "iters_run out of range");
  if (ws == nullptr || ws_bytes < plan_workspace(nullptr, 0, N, n, m, sizeof(T), true).total) {
    set_error("dicp_icp_backward: workspace too small");
    return DICP_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  IcpWorkspace w = plan_workspace(ws, ws_bytes, N, n, m, sizeof(T), true);
  const WeightParams P = make_weight_params(p);
  const bool fast = weight_params_fast(P);
  const TilePlan plan = bw_tile_plan(N, n);
  const int nseed = bw_seed_blocks(n), nrows = bw_rows_per_pair(N, n);
  const size_t rec_stride = (size_t)N * ST_STRIDE;
  const bool pl = p->icp_type == DICP_PT2PL;
  const int pair_blocks = (N + 3) / 4;         // one warp per pair
  auto aligned16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  // fast path: the streams move as 16-byte aligned slices; the gather table is rebuilt here (the workspace of the
  // forward call may have been reused in between), the gradient table lives in the workspace
  const bool staged = sizeof(T) == 4 && aligned16(source) && aligned16(saved_idx) && aligned16(w_init) && aligned16(grad_source) && aligned16(grad_w);

  if (staged) {
    int rc = gn_pack_rows<T>(target, N, m, S, pl, w.rows, st);
    if (rc) return rc;
    if (BW_GTAB) DICP_CUDA(cudaMemsetAsync(w.gtab, 0, (size_t)N * m * (pl ? 32 : 16), st));
    else DICP_CUDA(cudaMemsetAsync(grad_target, 0, (size_t)N * m * S * sizeof(T), st));
  } else {
    DICP_CUDA(cudaMemsetAsync(grad_target, 0, (size_t)N * m * S * sizeof(T), st));
  }
  if (grad_w != nullptr) DICP_CUDA(cudaMemsetAsync(grad_w, 0, (size_t)N * n * sizeof(T), st));
  bwd_seed_kernel<T><<<dim3(nseed, N), BW_THREADS, 0, st>>>(source, grad_pc, saved_state + (size_t)iters_run * rec_stride, n, grad_source, w.bw_partials, n_valid);
  DICP_LAUNCH_CHECK();
  constexpr size_t smem_rest = (BW_THREADS / 32) * 64 * sizeof(double) + 2 * BW_STAGES * sizeof(uint64_t);
  const bool unal = (n % 4) != 0;
  for (int k = iters_run - 1; k >= 0; --k) {
    const double* rec_k = saved_state + (size_t)k * rec_stride;
    const bool first = k == iters_run - 1;
    // the first reverse step folds the seed pass's rows (one per seed CTA, contiguous), later ones the point pass's
    if (first || !(BW_PDL & 1)) {
      bwd_pair_kernel<T><<<pair_blocks, 128, 0, st>>>(w.bw_partials, first ? nseed : nrows, first ? nseed : nrows, rec_k, p->dim, first ? 0 : 1,
                                                      grad_T, w.bw_state, grad_T_init, N);
      DICP_LAUNCH_CHECK();
    } else {     // scheduled while the point pass of iteration k + 1 drains (it waits for that grid's completion itself)
      DICP_CUDA(launch_pdl(bwd_pair_kernel<T>, dim3(pair_blocks), dim3(128), 0, st, (const double*)w.bw_partials, nrows, nrows, rec_k, p->dim, 1,
                           grad_T, w.bw_state, grad_T_init, N));
    }
    const int32_t* idx_k = saved_idx + (size_t)k * N * n;
    if (staged) {
      if constexpr (sizeof(T) == 4) {
#define DICP_BW_LAUNCH2(PL, FW, RG, UA)                                                                                            \
  do {                                                                                                                             \
    constexpr size_t smem = (size_t)BW_STAGES * (BW_TILE + ((UA) ? 8 : 0)) * 36 + smem_rest;                                       \
    if (smem > 48 * 1024)                                                                                                          \
      DICP_CUDA(cudaFuncSetAttribute(bwd_points_staged_kernel<PL, FW, RG, UA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    if (BW_PDL & 2) /* the prologue overlaps the pair kernel just launched */                                                      \
      DICP_CUDA(launch_pdl(bwd_points_staged_kernel<PL, FW, RG, UA>, dim3(plan.ctas), dim3(BW_THREADS), smem, st, source, (const void*)w.rows, w_init, \
                           idx_k, rec_k, (const double*)w.bw_state, n, m, P, grad_source, BW_GTAB ? (float*)w.gtab : (float*)grad_target, grad_w,  \
                           w.bw_partials, nrows, plan, S, n_valid));                                                               \
    else                                                                                                                           \
      bwd_points_staged_kernel<PL, FW, RG, UA><<<plan.ctas, BW_THREADS, smem, st>>>(source, w.rows, w_init, idx_k, rec_k, w.bw_state, n, m, P, \
                                                                         grad_source, BW_GTAB ? (float*)w.gtab : (float*)grad_target, grad_w, w.bw_partials, nrows, plan, S, n_valid); \
  } while (0)
#define DICP_BW_LAUNCH1(PL, FW, RG)              \
  do {                                           \
    if (unal) DICP_BW_LAUNCH2(PL, FW, RG, true); \
    else DICP_BW_LAUNCH2(PL, FW, RG, false);     \
  } while (0)
#define DICP_BW_LAUNCH(PL, FW)                         \
  do {                                                 \
    if (n_valid == nullptr) DICP_BW_LAUNCH1(PL, FW, false); \
    else DICP_BW_LAUNCH1(PL, FW, true);                \
  } while (0)
        if (pl && fast) DICP_BW_LAUNCH(true, true);
        else if (pl) DICP_BW_LAUNCH(true, false);
        else if (fast) DICP_BW_LAUNCH(false, true);
        else DICP_BW_LAUNCH(false, false);
#undef DICP_BW_LAUNCH
#undef DICP_BW_LAUNCH1
#undef DICP_BW_LAUNCH2
      }
    } else {
#define DICP_BW_LAUNCH(PL, FW)                                                                                                     \
  bwd_points_generic_kernel<T, PL, FW><<<plan.ctas, BW_THREADS, 0, st>>>(source, target, w_init, idx_k, rec_k, w.bw_state, n, m, S, P, \
                                                                        grad_source, grad_target, grad_w, w.bw_partials, nrows, plan, n_valid)
      if (pl && fast) DICP_BW_LAUNCH(true, true);
      else if (pl) DICP_BW_LAUNCH(true, false);
      else if (fast) DICP_BW_LAUNCH(false, true);
      else DICP_BW_LAUNCH(false, false);
#undef DICP_BW_LAUNCH
    }
    DICP_LAUNCH_CHECK();
  }
  bwd_pair_kernel<T><<<pair_blocks, 128, 0, st>>>(w.bw_partials, nrows, nrows, saved_state, p->dim, 2, grad_T, w.bw_state, grad_T_init, N);
  DICP_LAUNCH_CHECK();
  if (staged && BW_GTAB) {
    if constexpr (sizeof(T) == 4) {
      const size_t total = (size_t)N * m;
      const int grid = (int)((total + 255) / 256 < (size_t)device_sm_count() * 16 ? (total + 255) / 256 : (size_t)device_sm_count() * 16);
      if (pl) bwd_unpack_gtab_kernel<true><<<grid, 256, 0, st>>>((const float*)w.gtab, total, S, grad_target);
      else bwd_unpack_gtab_kernel<false><<<grid, 256, 0, st>>>((const float*)w.gtab, total, S, grad_target);
      DICP_LAUNCH_CHECK();
    }
  }
  return 0;
}

}  // namespace dicp

extern "C" {
int dicp_icp_backward_f32(const float* source, const float* target, const float* w_init, const int32_t* n_valid, int N, int n, int m, int S,
                          const dicp_params* params, int iters_run, const int32_t* saved_idx, const double* saved_state,
                          const float* grad_T, const float* grad_pc, float* grad_source, float* grad_target, float* grad_w,
                          float* grad_T_init, void* ws, size_t ws_bytes, void* stream) {
  return dicp::icp_backward_impl<float>(source, target, w_init, n_valid, N, n, m, S, params, iters_run, saved_idx, saved_state, grad_T, grad_pc,
                                        grad_source, grad_target, grad_w, grad_T_init, ws, ws_bytes, stream);
}
int dicp_icp_backward_f64(const double* source, const double* target, const double* w_init, const int32_t* n_valid, int N, int n, int m, int S,
                          const dicp_params* params, int iters_run, const int32_t* saved_idx, const double* saved_state,
                          const double* grad_T, const double* grad_pc, double* grad_source, double* grad_target, double* grad_w,
                          double* grad_T_init, void* ws, size_t ws_bytes, void* stream) {
  return dicp::icp_backward_impl<double>(source, target, w_init, n_valid, N, n, m, S, params, iters_run, saved_idx, saved_state, grad_T, grad_pc,
                                         grad_source, grad_target, grad_w, grad_T_init, ws, ws_bytes, stream);
}
}
