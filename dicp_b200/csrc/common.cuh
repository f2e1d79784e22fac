// common.cuh — internal helpers shared by the .cu files of libdicp_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/dicp_b200.h"
#include "dicp_math.cuh"

namespace dicp {

void set_error(const char* fmt, ...);

#define DICP_CHECK_ARG(cond, msg)                       \
  do {                                                  \
    if (!(cond)) {                                      \
      dicp::set_error("%s: %s", __func__, msg);         \
      return DICP_ERR_ARG;                              \
    }                                                   \
  } while (0)

#define DICP_CUDA(call)                                                            \
  do {                                                                             \
    cudaError_t _e = (call);                                                       \
    if (_e != cudaSuccess) {                                                       \
      dicp::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
      return (int)_e;                                                              \
    }                                                                              \
  } while (0)

#define DICP_LAUNCH_CHECK()                                                        \
  do {                                                                             \
    cudaError_t _e = cudaGetLastError();                                           \
    if (_e != cudaSuccess) {                                                       \
      dicp::set_error("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return (int)_e;                                                              \
    }                                                                              \
  } while (0)

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// bump allocator over the caller's workspace
struct Arena {
  char* base;
  size_t size, off;
  Arena(void* p, size_t n) : base((char*)p), size(n), off(0) {}
  template <typename U> U* take(size_t count) {
    size_t bytes = align_up(count * sizeof(U));
    if (base != nullptr && off + bytes > size) { off += bytes; return nullptr; }
    U* r = base ? (U*)(base + off) : nullptr;
    off += bytes;
    return r;
  }
};

inline WeightParams make_weight_params(const dicp_params* p) {
  WeightParams w;
  w.loss_id = p->loss_id;
  w.use_trim = p->use_trim;
  w.diff = p->differentiable;
  w.c = p->loss_metric;
  w.tau = p->trim_dist;
  w.steep = p->tanh_steepness;
  w.thresh = p->match_ratio_thresh;
  return w;
}

#if defined(__CUDACC__)
// gather of one target row (S = 3 | 6 columns, 8-byte aligned when S == 6 or the row index is even)
template <typename T, bool PT2PL>
__device__ __forceinline__ void load_target_row(const T* __restrict__ row, int S, T* q, T* nr) {
  if (PT2PL && sizeof(T) == 4) {            // 24-byte rows: three 8-byte loads
    const float2 a = *reinterpret_cast<const float2*>(row), b = *reinterpret_cast<const float2*>(row + 2),
                 c = *reinterpret_cast<const float2*>(row + 4);
    q[0] = (T)a.x; q[1] = (T)a.y; q[2] = (T)b.x; nr[0] = (T)b.y; nr[1] = (T)c.x; nr[2] = (T)c.y;
  } else {
    q[0] = row[0]; q[1] = row[1]; q[2] = row[2];
    if (PT2PL) { nr[0] = row[3]; nr[1] = row[4]; nr[2] = row[5]; }
    else { nr[0] = nr[1] = nr[2] = T(0); }
  }
}
// gather of one row of the re-laid fp32 target table (gn_pack_rows): ONE load instruction per
// correspondence -- 32-byte (x,y,z,nx,ny,nz,0,0) records through the 256-bit load of sm_100
// (LDG.E.256) for point-to-plane, 16-byte (x,y,z,0) records for point-to-point.  A warp's gather is
// fully divergent, so the L1 processes one wavefront per lane and load instruction: one
// instruction instead of three is a third of the L1 time, and the aligned record is one 32-byte
// sector where a 24-byte row straddles two of them half of the time.
template <typename T, bool PT2PL>
__device__ __forceinline__ void load_packed_row(const void* __restrict__ rows, size_t row, T* q, T* nr) {
  if constexpr (sizeof(T) == 4) {
    if constexpr (PT2PL) {
      const float* p = static_cast<const float*>(rows) + row * 8;
      float pad0, pad1;
      asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
          : "=f"(q[0]), "=f"(q[1]), "=f"(q[2]), "=f"(nr[0]), "=f"(nr[1]), "=f"(nr[2]), "=f"(pad0), "=f"(pad1)
          : "l"(p));
    } else {
      const float4 v = __ldg(static_cast<const float4*>(rows) + row);
      q[0] = v.x; q[1] = v.y; q[2] = v.z;
      nr[0] = nr[1] = nr[2] = 0.f;
    }
  }   // fp64 has no table: callers gather from the caller's rows (load_target_row)
}
// gather of one 24-byte fp32 target row (x,y,z,nx,ny,nz) from the caller's [.,6] tensor (16-byte aligned base, so a
// row starts at a multiple of 8 bytes): the two aligned 16-byte loads that cover the 32-byte window holding the
// row, then a select on the row's parity.  Against the 32-byte re-laid table this saves a quarter of the gather's
// DRAM traffic (the whole 24 B/row tensor is fetched once; the table is 32 B/row) and the table itself.
// `tensor_end`: the row is the last one of the whole tensor and the window would reach past it.
__device__ __forceinline__ void load_row24(const float* __restrict__ row, bool tensor_end, float* q, float* nr) {
  const bool odd = (reinterpret_cast<uintptr_t>(row) & 8) != 0;
  if (tensor_end && !odd) {                     // rare: plain 8-byte loads
    const float2 a = __ldg(reinterpret_cast<const float2*>(row)), b = __ldg(reinterpret_cast<const float2*>(row + 2)),
                 c = __ldg(reinterpret_cast<const float2*>(row + 4));
    q[0] = a.x; q[1] = a.y; q[2] = b.x; nr[0] = b.y; nr[1] = c.x; nr[2] = c.y;
    return;
  }
  const float4* w = reinterpret_cast<const float4*>(reinterpret_cast<uintptr_t>(row) & ~(uintptr_t)15);
  float4 v0, v1;
  asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v0.x), "=f"(v0.y), "=f"(v0.z), "=f"(v0.w) : "l"(w));
  asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v1.x), "=f"(v1.y), "=f"(v1.z), "=f"(v1.w) : "l"(w + 1));
  q[0] = odd ? v0.z : v0.x; q[1] = odd ? v0.w : v0.y; q[2] = odd ? v1.x : v0.z;
  nr[0] = odd ? v1.y : v0.w; nr[1] = odd ? v1.z : v1.x; nr[2] = odd ? v1.w : v1.y;
}
// scatter-add of one 6-column gradient row: fp32 uses the 8-byte vector reduction (red.global.add.v2.f32)
__device__ __forceinline__ void atomic_add_row6(float* row, const float* v) {
  atomicAdd(reinterpret_cast<float2*>(row), make_float2(v[0], v[1]));
  atomicAdd(reinterpret_cast<float2*>(row + 2), make_float2(v[2], v[3]));
  atomicAdd(reinterpret_cast<float2*>(row + 4), make_float2(v[4], v[5]));
}
__device__ __forceinline__ void atomic_add_row6(double* row, const double* v) {
#pragma unroll
  for (int c = 0; c < 6; ++c) atomicAdd(row + c, v[c]);
}
#endif

// device control block of one forward call
struct Ctrl {
  int stop;        // all pairs below tolerance (only raised when !const_iter)
  int iters_run;   // iterations actually executed
  int singular;    // a normal matrix was exactly singular
  int ticket;      // per-iteration arrival counter of the update kernel
  int nbelow;      // per-iteration count of pairs below tolerance
};

// ---- nearest neighbour (nn_kernels.cu) ------------------------------------------------
size_t nn_workspace_bytes(int N, int n, int m, int elem_size);
// packed targets: [N, m_pad] records, m_pad = round_up(m, NN_CHUNK), sentinel rows.
//   float : [N,m_pad] filter records (-2y', |y'|^2) in the pair's translated frame (16 B), then [N,m_pad]
//           original (x,y,z,0) records for the exact refine / fallback, then [N] centres
//   double: [N,m_pad] (x,y,z,0), 32 B, direct-form scan
template <typename T> struct Vec4;
template <> struct Vec4<float> { typedef float4 type; };
template <> struct Vec4<double> { typedef double4 type; };
constexpr int NN_CHUNK = 32;
inline int nn_m_pad(int m) { return (m + NN_CHUNK - 1) / NN_CHUNK * NN_CHUNK; }
size_t nn_packed_bytes(int N, int m, int elem_size);     // nn_kernels.cu (depends on the fp32 record layout chosen there)

template <typename T>
int nn_pack_targets(const T* tgt, int N, int m, int S, const int32_t* m_valid, typename Vec4<T>::type* packed, cudaStream_t st);
// state: nullptr (queries are used as given) or pointer to this iteration's saved-state
// records [N, ST_STRIDE] whose C,r transform the queries first (ICP.py:137).
template <typename T>
int nn_search(const T* q, const double* state, const typename Vec4<T>::type* packed, int N, int n, int m,
              int32_t* idx_out, T* d2_out, const int* stop_flag, void* ws, size_t ws_bytes, cudaStream_t st,
              const int32_t* n_valid = nullptr);   // n_valid [N] or NULL: real query count per pair; rows beyond it are padding
size_t nn_search_ws_bytes(int N, int n, int m, int elem_size);

// ---- uniform-grid exact NN (grid_nn_kernels.cu) ---------------------------------------
struct GridWorkspace {
  double* gparams;      // [N, 8]  origin xyz, h, 1/h
  int* gdims;           // [N, 4]  nx, ny, nz, ncells
  int* cell_start;      // [N, cells_stride]
  int* cell_count;      // [N, cells_stride]  histogram, then scatter cursor
  int* cell_of;         // [N, m]
  void* recs;           // [N, m] GridRec<T>
  int* max_occupancy;   // [2]
  int* q_cell_start;    // [N, cells_stride]  query sort
  int* q_cell_of;       // [N, n]
  int* q_order;         // [N, n]  query indices in cell order
  int cells_stride;
  size_t total;
};
GridWorkspace grid_ws_layout(void* ws, size_t ws_bytes, int N, int n, int m, int elem_size);
template <typename T> int grid_build(const T* tgt, int N, int m, int S, const int32_t* m_valid, const GridWorkspace& g, cudaStream_t st);
template <typename T> int grid_sort_queries(const T* q, const double* state, int N, int n, const GridWorkspace& g, cudaStream_t st);
template <typename T>
int grid_search(const T* q, const double* state, int N, int n, int m, const GridWorkspace& g, bool sorted, int32_t* idx_out, T* d2_out,
                const int* stop_flag, cudaStream_t st, const T* tgt = nullptr, int S = 3, const int32_t* warm_idx = nullptr);
// warm_idx [N,n] (with tgt [N,m,S]): a neighbour candidate per query, e.g. the previous iteration's, that bounds the search

// ---- Gauss-Newton (gn_kernels.cu) -----------------------------------------------------
int gn_blocks_per_pair(int N, int n);   // partial-sum rows per pair of the accumulate kernel's tile plan (tile_stream.cuh)
int bw_rows_per_pair(int N, int n);      // rows per pair of the backward's [N, rows, BW_NRED] partial-sum table
size_t bw_gtab_bytes(int N, int m, int elem_size);   // fp32: aligned gradient table of the targets (8 floats per row)
// where the accumulate kernel finds (C, r, active) of pair b: state + b * stride + {c, r, active}
// (active < 0: every pair is active).  The loop passes its saved-state records, the stage API a bare [N,12].
struct StateView { const double* base; int stride, c, r, active; };
inline StateView state_view_records(const double* rec) { return StateView{rec, ST_STRIDE, ST_C, ST_R, ST_ACTIVE}; }
// fp32 gather table of the targets (see load_packed_row); 0 bytes / no-op for fp64
size_t gn_rows_bytes(int N, int m, int elem_size);
template <typename T> int gn_pack_rows(const T* target, int N, int m, int S, bool pt2pl, void* rows, cudaStream_t st);
template <typename T>
int gn_accumulate(const T* source, const T* target, const void* rows, const T* w_init, const int32_t* idx, StateView state,
                  int N, int n, int m, int S, const dicp_params* p, T* weights_out, long long w_pair_stride,
                  double* partials, const int* stop_flag, cudaStream_t st, const int32_t* n_valid = nullptr);


// ---- workspace of one forward / backward call (gn_kernels.cu) -------------------------
constexpr int BW_STRIDE = 32;   // doubles per pair of backward carry: Cbar 9, rbar 3, lambda 6, delta 6, active 1
constexpr int BW_CBAR = 0, BW_RBAR = 9, BW_LAM = 12, BW_DEL = 18, BW_ACTIVE = 24;
constexpr int BW_NRED = 12;     // per-CTA partial sums of the backward point pass: Cbar 9, rbar 3
struct IcpWorkspace {
  void* packed;        // [N, m_pad] Vec4
  void* nn_ws;         // split partials of the NN search
  size_t nn_ws_bytes;
  double* partials;    // [N, nblk, RED_N]
  Ctrl* ctrl;
  double* bw_state;    // [N, BW_STRIDE]
  double* bw_partials; // [N, nblk, BW_NRED]
  void* grid_ws;       // uniform-grid index of the targets (nn_mode == DICP_NN_GRID)
  size_t grid_ws_bytes;
  void* rows;          // fp32: [N, m] aligned gather records of the targets (gn_pack_rows); unused for fp64
  void* gtab;          // fp32: [N, m] aligned gradient records of the targets (backward scatter); unused for fp64
  size_t total;
};
IcpWorkspace plan_workspace(void* ws, size_t ws_bytes, int N, int n, int m, int elem_size, bool backward_only);

}  // namespace dicp
