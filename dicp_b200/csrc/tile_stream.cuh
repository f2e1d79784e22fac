// tile_stream.cuh — what the two HBM-bound per-correspondence kernels (Gauss-Newton accumulate, backward
// point pass) share: a persistent-CTA tile schedule with deterministic partial-sum rows, and a shared-memory
// staging ring filled by TMA bulk copies (cp.async.bulk + mbarrier complete_tx) for the contiguous streams.
//
// The correspondences of a batch are N*n independent items whose per-pair sums must be reduced; tile t = b * tpp + j
// is the j-th TILE-sized slice of pair b.  The grid is exactly sm_count x CTAs-per-SM persistent CTAs (no tail
// wave).  The schedule runs in WAVES: in wave w CTA c owns the chunk of k consecutive tiles starting at
// (w G + c) k, so at any moment the G CTAs work on ~G k consecutive tiles, i.e. on G k / tpp neighbouring pairs.
// That locality is what keeps the random 32-byte row gathers cheap: DRAM is read in 64-byte granules, and the
// other half of a granule is only useful if the thread that needs it runs while the granule is still in L2
// (measured: with one contiguous tile range per CTA, all pairs are live at once, the 134 MB gather table of the
// C3 batch no longer fits the 126 MB L2 and DRAM reads go from 201 MB to 333 MB per launch).  The tiles left
// after the last full wave are split evenly in one final wave of smaller chunks (k_last), so no CTA idles for
// more than one tile.  A chunk may straddle a pair boundary; every (chunk, pair) segment reduces its sums once
// and writes one row of the pair's partial-sum table; rows are numbered by the chunk boundaries inside the pair,
// which makes the layout a pure function of the plan (deterministic summation order in the consumer).
//
// The contiguous inputs of a tile (indices, source points, prior weights, running gradients) arrive in shared
// memory through bulk copies issued by one thread several tiles ahead, so the only loads that occupy registers
// while in flight are the gathered target rows.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dicp {

struct TilePlan {
  int tile;            // correspondences per tile
  int tpp;             // tiles per pair = ceil(n / tile)
  int ctas;            // persistent CTAs G
  int k;               // tiles per chunk in the full waves
  int k_last;          // tiles per chunk in the final wave (0: none)
  int waves_full;      // number of full waves
  long long tiles;     // T = N * tpp
  long long full;      // F = waves_full * G * k tiles are covered by the full waves
  int adv_b, adv_j;    // a CTA's chunk starts advance by G * k tiles per wave = adv_b pairs + adv_j tiles
  int rows;            // partial-sum rows reserved per pair
};

// number of chunk starts in [0, x]
#if defined(__CUDACC__)
#define DICP_TS_HD __host__ __device__ __forceinline__
#else
#define DICP_TS_HD inline
#endif
DICP_TS_HD long long tp_bounds(const TilePlan& p, long long x) {
  return x < p.full ? x / p.k + 1 : p.full / p.k + (x - p.full) / p.k_last + 1;
}

// number of SMs of the current device (148 on B200); cached, 148 when no device is present (size queries on a
// GPU-less host must still answer, and they only have to agree with the launches of the same process)
int device_sm_count();

inline TilePlan make_tile_plan(int N, int n, int tile, int ctas_per_sm, int k) {
  TilePlan p;
  p.tile = tile;
  p.tpp = (n + tile - 1) / tile;
  p.tiles = (long long)N * p.tpp;
  long long g = (long long)device_sm_count() * ctas_per_sm;
  if (g > p.tiles) g = p.tiles;
  if (g < 1) g = 1;
  p.ctas = (int)g;
  p.k = k;
  p.waves_full = (int)(p.tiles / (g * k));
  p.full = (long long)p.waves_full * g * k;
  p.adv_b = (int)((g * k) / p.tpp);
  p.adv_j = (int)((g * k) % p.tpp);
  const long long rest = p.tiles - p.full;
  p.k_last = rest > 0 ? (int)((rest + g - 1) / g) : 0;
  const int kmin = p.waves_full > 0 ? (p.k_last > 0 && p.k_last < k ? p.k_last : k) : p.k_last;
  long long rows = (p.tpp - 1 + kmin - 1) / kmin + 1;           // chunk boundaries strictly inside a pair, plus one
  p.rows = (int)rows;
  return p;
}

// Walks the tiles of one CTA in schedule order without a division per tile or per chunk: the chunk starts of the
// full waves advance by a constant (adv_b pairs, adv_j tiles), the one chunk of the final wave is located once per
// CTA (TileWalk).  b / j: pair and tile-in-pair of the current tile; L: how many tiles this CTA has walked before
// it; seg_first: the tile starts a (chunk, pair) segment.
struct TileWalk {
  int cta;
  int last_b, last_j, last_len;      // this CTA's chunk of the final wave (last_len 0: none)
  int ntiles;                        // tiles the CTA walks in total
  DICP_TS_HD void init(const TilePlan& p, int c) {
    cta = c;
    last_b = last_j = last_len = 0;
    if (p.k_last > 0) {
      const long long s = p.full + (long long)c * p.k_last;
      const long long left = p.tiles - s;
      if (left > 0) {
        last_len = left < p.k_last ? (int)left : p.k_last;
        last_b = (int)(s / p.tpp);
        last_j = (int)(s - (long long)last_b * p.tpp);
      }
    }
    ntiles = p.waves_full * p.k + last_len;
  }
};
struct TileCursor {
  int b, j, L, rem, wave, sb, sj;
  bool seg_first;      // the current tile starts a (chunk, pair) segment
  DICP_TS_HD void init(const TilePlan& p, const TileWalk& w) {
    L = 0; wave = 0; seg_first = true;
    if (p.waves_full > 0) {
      const long long s = (long long)w.cta * p.k;
      sb = (int)(s / p.tpp);
      sj = (int)(s - (long long)sb * p.tpp);
      rem = p.k;
    } else {
      sb = w.last_b; sj = w.last_j; rem = w.last_len;
    }
    b = sb; j = sj;
  }
  DICP_TS_HD void next(const TilePlan& p, const TileWalk& w) {
    ++L;
    if (--rem == 0) {
      ++wave;
      seg_first = true;
      if (wave < p.waves_full) {
        sj += p.adv_j; sb += p.adv_b;
        if (sj >= p.tpp) { sj -= p.tpp; ++sb; }
        rem = p.k;
      } else {
        sb = w.last_b; sj = w.last_j; rem = w.last_len;       // past the end this is never looked at (callers stop at ntiles)
      }
      b = sb; j = sj;
      return;
    }
    seg_first = false;
    if (++j == p.tpp) { j = 0; ++b; seg_first = true; }
  }
  DICP_TS_HD long long tile(const TilePlan& p) const { return (long long)b * p.tpp + j; }
};

// row of pair b's partial-sum table that the segment starting at tile seg_tile writes
DICP_TS_HD int tp_segment_row(const TilePlan& p, int b, long long seg_tile) {
  return (int)(tp_bounds(p, seg_tile) - tp_bounds(p, (long long)b * p.tpp));
}

#if defined(__CUDACC__)
// Sum of NV (<= 32) per-lane values over the 32 lanes of a warp with 31 shuffles instead of 5 NV: recursive
// halving -- at the step with lane mask s the lanes whose bit s is set keep the upper half of the values and
// hand the lower half to their partner.  On return lane l holds in v[0] the warp total of value l (fixed tree).
template <typename T, int NV>
__device__ __forceinline__ void warp_multi_reduce(T (&v)[32]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int i = NV; i < 32; ++i) v[i] = T(0);
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const T send = up ? v[i] : v[i + s];
      const T keep = up ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Programmatic dependent launch (PDL).  A kernel launched with launch_pdl() may begin while its predecessor in the
// stream is still running, as soon as that one has executed pdl_trigger() in every CTA (or exited); everything the
// predecessor wrote is visible after pdl_wait(), which returns when the predecessor grid has completed.  Both are
// no-ops in a kernel that was launched the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
// asynchronous prefetch of a contiguous global range into L2 (no destination, no completion to wait for):
// src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
// L2 eviction-priority policies for cache hints, and the hinted forms of the bulk copy / prefetch
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2_hint(const void* src, uint32_t bytes, uint64_t policy) {
  asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(src), "r"(bytes), "l"(policy) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// per-thread asynchronous 8-byte copy global -> shared (LDGSTS), completed by cp_async_wait_all
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// TMA bulk copy global -> shared (UBLKCP): dst / src 16-byte aligned, bytes a multiple of 16; completion is
// signalled on `bar` as transaction bytes
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// host side: the launch attribute that lets `kernel` overlap the tail of its predecessor (see pdl_wait)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

}  // namespace dicp
