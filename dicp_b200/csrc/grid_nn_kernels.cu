// grid_nn_kernels.cu — exact nearest neighbour through a per-pair uniform grid.
//
// The brute-force scan (nn_kernels.cu) evaluates n*m distances per pair and iteration; the
// targets of a call are static across the iterations (ICP.py:132-260 only moves the source), so
// a spatial index built once per call pays for itself in the first iteration.  Per pair:
//   grid_bbox_kernel     bounding box of the targets -> cell size h = cbrt(volume / m) (about one
//                        target per cell of the occupied slab), dims <= cells_cap
//   grid_count_kernel    histogram of targets per cell (atomicAdd)
//   grid_scan_kernel     exclusive prefix sum per pair (one CTA per pair) -> cell_start
//   grid_scatter_kernel  counting-sort scatter of (x,y,z,index) records into cell order
//   grid_q*_kernel       counting sort of the queries by the cell of their initial position (once per
//                        call), so that the lanes of a warp search neighbouring cells
//   grid_query_kernel    one thread per query: visit the 3x3x3 block around the query's cell, then
//                        ring by ring (Chebyshev distance 2,3,..), direct-form fp32/fp64 distance to every
//                        record, lexicographic (d2, index) minimum; stop as soon as the best
//                        distance is strictly smaller than the distance from the query to the
//                        faces of the visited block behind which unvisited cells remain.
// The result is the same argmin the brute-force search returns (same direct-form arithmetic,
// same first-minimum tie-break); tests compare the two bit for bit.  Memory- and latency-bound
// (scattered 16-byte record reads from L2), O(n) per pair instead of O(n*m).
#include <stdlib.h>

#include "common.cuh"

namespace dicp {

// per-pair grid parameters (doubles): origin xyz, cell size h, 1/h; the dims live in gdims[b] = {nx, ny, nz, ncells}
constexpr int GP_ORIGIN = 0;
constexpr int GP_H = 3;
constexpr int GP_INVH = 4;
constexpr int GP_STRIDE = 8;

template <typename T> struct GridRec { T x, y, z; int idx; };   // 16 B (fp32) / 32 B (fp64): one load per candidate

template <typename T>
__global__ void grid_bbox_kernel(const T* __restrict__ tgt, int m_all, int S, const int32_t* __restrict__ m_valid, int cells_cap,
                                 double hscale, double* __restrict__ gparams, int* __restrict__ gdims) {
  __shared__ double smin[3][8], smax[3][8];
  const int b = blockIdx.x;
  const T* tb = tgt + (size_t)b * m_all * S;
  const int m = m_valid != nullptr ? min(max(m_valid[b], 1), m_all) : m_all;
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int j = threadIdx.x; j < m; j += blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double v = (double)tb[(size_t)j * S + c];
      lo[c] = fmin(lo[c], v);
      hi[c] = fmax(hi[c], v);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[c] = fmin(lo[c], __shfl_down_sync(0xffffffffu, lo[c], o));
      hi[c] = fmax(hi[c], __shfl_down_sync(0xffffffffu, hi[c], o));
    }
    if (lane == 0) { smin[c][warp] = lo[c]; smax[c][warp] = hi[c]; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ext[3];
    for (int c = 0; c < 3; ++c) {
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { smin[c][0] = fmin(smin[c][0], smin[c][w]); smax[c][0] = fmax(smax[c][0], smax[c][w]); }
      ext[c] = smax[c][0] - smin[c][0];
    }
    const double emax = fmax(ext[0], fmax(ext[1], ext[2]));
    const double floor_ext = fmax(emax * 1e-3, 1e-30);               // flat / degenerate clouds: thin slab, not zero volume
    double vol = 1.0;
    for (int c = 0; c < 3; ++c) vol *= fmax(ext[c], floor_ext);
    double h = hscale * cbrt(vol / (double)m);
    if (!(h > 0.0)) h = 1.0;
    int d[3];
    for (int it = 0; it < 64; ++it) {
      long long prod = 1;
      for (int c = 0; c < 3; ++c) {
        double q = floor(ext[c] / h) + 1.0;
        d[c] = q > 1024.0 ? 1024 : (int)q;
        if (ext[c] / h >= 1024.0) prod = (long long)cells_cap + 1;     // too fine for this axis
        prod *= d[c];
      }
      if (prod <= cells_cap) break;
      h *= 1.26;
    }
    double* gp = gparams + (size_t)b * GP_STRIDE;
    for (int c = 0; c < 3; ++c) gp[GP_ORIGIN + c] = smin[c][0];
    gp[GP_H] = h;
    gp[GP_INVH] = 1.0 / h;
    int* gd = gdims + (size_t)b * 4;
    gd[0] = d[0]; gd[1] = d[1]; gd[2] = d[2]; gd[3] = d[0] * d[1] * d[2];
  }
}

__device__ __forceinline__ int cell_coord(double v, double origin, double invh, int dim) {
  int c = (int)floor((v - origin) * invh);
  return c < 0 ? 0 : (c >= dim ? dim - 1 : c);
}

template <typename T>
__global__ void grid_count_kernel(const T* __restrict__ tgt, int m, int S, const int32_t* __restrict__ m_valid, int cells_stride,
                                  const double* __restrict__ gparams, const int* __restrict__ gdims, int* __restrict__ cell_count,
                                  int* __restrict__ cell_of) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= (m_valid != nullptr ? min(max(m_valid[b], 1), m) : m)) return;
  const double* gp = gparams + (size_t)b * GP_STRIDE;
  const int* gd = gdims + (size_t)b * 4;
  const T* row = tgt + ((size_t)b * m + j) * S;
  const int cx = cell_coord((double)row[0], gp[GP_ORIGIN], gp[GP_INVH], gd[0]);
  const int cy = cell_coord((double)row[1], gp[GP_ORIGIN + 1], gp[GP_INVH], gd[1]);
  const int cz = cell_coord((double)row[2], gp[GP_ORIGIN + 2], gp[GP_INVH], gd[2]);
  const int cell = (cz * gd[1] + cy) * gd[0] + cx;
  cell_of[(size_t)b * m + j] = cell;
  atomicAdd(&cell_count[(size_t)b * cells_stride + cell], 1);
}

// exclusive scan of cell_count -> cell_start (cells_stride entries per pair, last valid = ncells), and the
// maximum cell occupancy of the pair
__global__ void grid_scan_kernel(const int* __restrict__ cell_count, int cells_stride, const int* __restrict__ gdims,
                                 int* __restrict__ cell_start, int* __restrict__ max_occupancy) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  const int b = blockIdx.x;
  const int ncells = gdims[(size_t)b * 4 + 3];
  const int* cnt = cell_count + (size_t)b * cells_stride;
  int* st = cell_start + (size_t)b * cells_stride;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  int mx = 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int base = 0; base < ncells; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = i < ncells ? cnt[i] : 0;
    mx = max(mx, v);
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int t = lane < nwarps ? warp_tot[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      warp_tot[lane] = t;                                  // inclusive totals of the warps
    }
    __syncthreads();
    const int before = carry + (warp > 0 ? warp_tot[warp - 1] : 0) + (x - v);
    if (i < ncells) st[i] = before;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) st[ncells] = carry;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_down_sync(0xffffffffu, mx, o));
  if (lane == 0) atomicMax(max_occupancy, mx);
}

template <typename T>
__global__ void grid_scatter_kernel(const T* __restrict__ tgt, int m, int S, const int32_t* __restrict__ m_valid, int cells_stride,
                                    const int* __restrict__ cell_of, const int* __restrict__ cell_start, int* __restrict__ cell_cursor,
                                    GridRec<T>* __restrict__ recs) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= (m_valid != nullptr ? min(max(m_valid[b], 1), m) : m)) return;
  const int cell = cell_of[(size_t)b * m + j];
  const int pos = cell_start[(size_t)b * cells_stride + cell] + atomicAdd(&cell_cursor[(size_t)b * cells_stride + cell], 1);
  const T* row = tgt + ((size_t)b * m + j) * S;
  GridRec<T> r;
  r.x = row[0]; r.y = row[1]; r.z = row[2]; r.idx = j;
  recs[(size_t)b * m + pos] = r;
}

// visit the records of cells [c0, c1] of one grid row (contiguous in the sorted record array)
template <typename T>
__device__ __forceinline__ void grid_visit_run(const int* __restrict__ cs, const GridRec<T>* __restrict__ rb, int c0, int c1,
                                               T x0, T x1, T x2, T& best, int& bi) {
  const int s = cs[c0], e = cs[c1 + 1];
  for (int k = s; k < e; ++k) {
    const GridRec<T> t = rb[k];
    const T d = dist2<T>(x0, x1, x2, t.x, t.y, t.z);
    if (d < best || (d == best && t.idx < bi)) { best = d; bi = t.idx; }      // lexicographic: first minimum
  }
}

template <typename T>
__global__ void __launch_bounds__(128)
grid_query_kernel(const T* __restrict__ q, const double* __restrict__ state, int n, int m, int cells_stride,
                  const double* __restrict__ gparams, const int* __restrict__ gdims, const int* __restrict__ cell_start,
                  const GridRec<T>* __restrict__ recs, const int* __restrict__ order, int32_t* __restrict__ idx_out,
                  T* __restrict__ d2_out, const int* __restrict__ stop_flag, const T* __restrict__ tgt, int S,
                  const int32_t* __restrict__ warm_idx) {
  if (stop_flag != nullptr && *stop_flag != 0) return;
  const int b = blockIdx.y;
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= n) return;
  // queries are processed in the cell order of their initial position, so the lanes of a warp walk
  // the same few cells (coalesced / broadcast record reads, similar trip counts)
  const int i = order != nullptr ? order[(size_t)b * n + slot] : slot;
  const T* p = q + ((size_t)b * n + i) * 3;
  T x0 = p[0], x1 = p[1], x2 = p[2];
  if (state != nullptr) {
    const double* rec = state + (size_t)b * ST_STRIDE;
    T C[9], pv[3] = {x0, x1, x2}, a[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) C[k] = (T)rec[ST_C + k];
    rotate(C, pv, a);                                                  // ICP.py:137
    x0 = a[0] + (T)rec[ST_R]; x1 = a[1] + (T)rec[ST_R + 1]; x2 = a[2] + (T)rec[ST_R + 2];
  }
  const double* gp = gparams + (size_t)b * GP_STRIDE;
  const int* gd = gdims + (size_t)b * 4;
  const int nx = gd[0], ny = gd[1], nz = gd[2];
  const double h = gp[GP_H], ox = gp[GP_ORIGIN], oy = gp[GP_ORIGIN + 1], oz = gp[GP_ORIGIN + 2];
  const int cx = cell_coord((double)x0, ox, gp[GP_INVH], nx);
  const int cy = cell_coord((double)x1, oy, gp[GP_INVH], ny);
  const int cz = cell_coord((double)x2, oz, gp[GP_INVH], nz);
  const int* cs = cell_start + (size_t)b * cells_stride;
  const GridRec<T>* rb = recs + (size_t)b * m;

  T best = sizeof(T) == 4 ? (T)__int_as_float(0x7f800000) : (T)__longlong_as_double(0x7ff0000000000000LL);
  int bi = 0x7fffffff;
  const size_t o = (size_t)b * n + i;
  if (warm_idx != nullptr) {
    // warm start (the solver loop from its second iteration on): the neighbour of the previous iteration is a real
    // candidate, so its distance bounds the search from the start; the result is still the lexicographic minimum
    const int jw = warm_idx[o];
    const T* y = tgt + ((size_t)b * m + jw) * S;
    best = dist2<T>(x0, x1, x2, y[0], y[1], y[2]);
    bi = jw;
  }
  const int rmax = max(max(max(cx, nx - 1 - cx), max(cy, ny - 1 - cy)), max(cz, nz - 1 - cz));
  constexpr int R0 = 1;
  const int m_recs = cs[gd[3]];               // records of this pair (= valid targets)
  {
    // ring 1 = the 3x3x3 block around the query's cell, as 9 row runs along x, nearest rows first.  A row, or an end
    // cell of a row, is skipped when every point of it is further away than the best target so far: squared distance
    // to the faces that separate it from the query's cell (0 towards a side the query is not inside of), with the
    // same slack for the fp32 rounding of `best` as the termination rule below; strict, so equal distances are seen
    auto sq0 = [](double v) { return v > 0.0 ? v * v : 0.0; };
    const double lx[3] = {sq0((double)x0 - (ox + cx * h)), 0.0, sq0((ox + (cx + 1) * h) - (double)x0)};
    const double ly[3] = {sq0((double)x1 - (oy + cy * h)), 0.0, sq0((oy + (cy + 1) * h) - (double)x1)};
    const double lz[3] = {sq0((double)x2 - (oz + cz * h)), 0.0, sq0((oz + (cz + 1) * h) - (double)x2)};
    constexpr int ROWS[9][2] = {{0, 0}, {-1, 0}, {1, 0}, {0, -1}, {0, 1}, {-1, -1}, {1, -1}, {-1, 1}, {1, 1}};
    constexpr double SLACK = 1.0 - 1e-6;
#pragma unroll
    for (int w = 0; w < 9; ++w) {
      const int dy = ROWS[w][0], dz = ROWS[w][1];
      const int y = cy + dy, z = cz + dz;
      if (y < 0 || y > ny - 1 || z < 0 || z > nz - 1) continue;
      const double lb2 = ly[dy + 1] + lz[dz + 1];
      if ((double)best < lb2 * SLACK) continue;
      const int rowbase = (z * ny + y) * nx;
      const int c0 = (cx - 1 >= 0 && !((double)best < (lb2 + lx[0]) * SLACK)) ? cx - 1 : cx;
      const int c1 = (cx + 1 <= nx - 1 && !((double)best < (lb2 + lx[2]) * SLACK)) ? cx + 1 : cx;
      grid_visit_run<T>(cs, rb, rowbase + c0, rowbase + c1, x0, x1, x2, best, bi);
    }
  }
  for (int r = R0; r <= max(rmax, R0); ++r) {
    if (r > R0) {
      // a query far from every target (bad initial guess, outlier) would walk O(r^3) empty cells: once
      // the rows of the next ring outnumber the records, scan the records linearly instead (bounded work)
      if ((long long)(2 * r + 1) * (2 * r + 1) * 4 > (long long)max(m_recs, 256)) {
        grid_visit_run<T>(cs, rb, 0, gd[3] - 1, x0, x1, x2, best, bi);
        break;
      }
      const int z0 = max(cz - r, 0), z1 = min(cz + r, nz - 1);
      const int y0 = max(cy - r, 0), y1 = min(cy + r, ny - 1);
      const int xa = max(cx - r, 0), xb = min(cx + r, nx - 1);
      for (int z = z0; z <= z1; ++z) {
        const bool zface = (z == cz - r) || (z == cz + r);
        for (int y = y0; y <= y1; ++y) {
          const int rowbase = (z * ny + y) * nx;
          if (zface || (y == cy - r) || (y == cy + r)) {
            grid_visit_run<T>(cs, rb, rowbase + xa, rowbase + xb, x0, x1, x2, best, bi);      // whole x-run is on the ring
          } else {                                                                              // interior row: the two end cells
            if (cx - r >= 0) grid_visit_run<T>(cs, rb, rowbase + cx - r, rowbase + cx - r, x0, x1, x2, best, bi);
            if (cx + r <= nx - 1) grid_visit_run<T>(cs, rb, rowbase + cx + r, rowbase + cx + r, x0, x1, x2, best, bi);
          }
        }
      }
    }
    // every unvisited target lies behind one of the faces of the visited block [c-r, c+r] that still
    // has cells behind it (or in a part of ring 1 that was skipped because it cannot hold anything closer);
    // its distance is at least the query's distance to that face
    double lb = 1e300;
    if (cx - r > 0) lb = fmin(lb, (double)x0 - (ox + (cx - r) * h));
    if (cx + r < nx - 1) lb = fmin(lb, (ox + (cx + r + 1) * h) - (double)x0);
    if (cy - r > 0) lb = fmin(lb, (double)x1 - (oy + (cy - r) * h));
    if (cy + r < ny - 1) lb = fmin(lb, (oy + (cy + r + 1) * h) - (double)x1);
    if (cz - r > 0) lb = fmin(lb, (double)x2 - (oz + (cz - r) * h));
    if (cz + r < nz - 1) lb = fmin(lb, (oz + (cz + r + 1) * h) - (double)x2);
    if (lb > 0.0 && (double)best < lb * lb * (1.0 - 1e-6)) break;      // strict: an equal-distance target cannot hide behind
  }
  idx_out[o] = bi == 0x7fffffff ? 0 : bi;
  if (d2_out != nullptr) d2_out[o] = best;
}

// counting sort of the queries by the grid cell of their (transformed) position -> order[b, 0..n)
template <typename T>
__global__ void grid_qcount_kernel(const T* __restrict__ q, const double* __restrict__ state, int n, int cells_stride,
                                   const double* __restrict__ gparams, const int* __restrict__ gdims, int* __restrict__ cell_count,
                                   int* __restrict__ cell_of) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const T* p = q + ((size_t)b * n + i) * 3;
  T x0 = p[0], x1 = p[1], x2 = p[2];
  if (state != nullptr) {
    const double* rec = state + (size_t)b * ST_STRIDE;
    T C[9], pv[3] = {x0, x1, x2}, a[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) C[k] = (T)rec[ST_C + k];
    rotate(C, pv, a);
    x0 = a[0] + (T)rec[ST_R]; x1 = a[1] + (T)rec[ST_R + 1]; x2 = a[2] + (T)rec[ST_R + 2];
  }
  const double* gp = gparams + (size_t)b * GP_STRIDE;
  const int* gd = gdims + (size_t)b * 4;
  const int cx = cell_coord((double)x0, gp[GP_ORIGIN], gp[GP_INVH], gd[0]);
  const int cy = cell_coord((double)x1, gp[GP_ORIGIN + 1], gp[GP_INVH], gd[1]);
  const int cz = cell_coord((double)x2, gp[GP_ORIGIN + 2], gp[GP_INVH], gd[2]);
  const int cell = (cz * gd[1] + cy) * gd[0] + cx;
  cell_of[(size_t)b * n + i] = cell;
  atomicAdd(&cell_count[(size_t)b * cells_stride + cell], 1);
}
__global__ void grid_qscatter_kernel(int n, int cells_stride, const int* __restrict__ cell_of, const int* __restrict__ cell_start,
                                     int* __restrict__ cell_cursor, int* __restrict__ order) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int cell = cell_of[(size_t)b * n + i];
  const int pos = cell_start[(size_t)b * cells_stride + cell] + atomicAdd(&cell_cursor[(size_t)b * cells_stride + cell], 1);
  order[(size_t)b * n + pos] = i;
}

#ifndef GRID_HSCALE_V
#define GRID_HSCALE_V 0.75    // with the row pruning of grid_query_kernel: 0.6 -> 5.02 ms per C3 step, 0.75 -> 4.95, 0.9 -> 5.03, 1.1 -> 5.17
#endif
constexpr double GRID_HSCALE = GRID_HSCALE_V;   // cell size in units of cbrt(box volume / m)

// ---- host side --------------------------------------------------------------------------------
int grid_cells_cap(int m) {
  long long c = 8LL * m;
  if (c < 64) c = 64;
  if (c > (1LL << 22)) c = 1LL << 22;
  return (int)c;
}

GridWorkspace grid_ws_layout(void* ws, size_t ws_bytes, int N, int n, int m, int elem_size) {
  Arena ar(ws, ws_bytes);
  GridWorkspace g;
  g.cells_stride = grid_cells_cap(m) + 1;
  g.gparams = ar.take<double>((size_t)N * GP_STRIDE);
  g.gdims = ar.take<int>((size_t)N * 4);
  g.cell_start = ar.take<int>((size_t)N * g.cells_stride);
  g.cell_count = ar.take<int>((size_t)N * g.cells_stride);      // reused as the scatter cursor
  g.cell_of = ar.take<int>((size_t)N * m);
  g.recs = ar.take<char>((size_t)N * m * (elem_size == 4 ? sizeof(GridRec<float>) : sizeof(GridRec<double>)));
  g.max_occupancy = ar.take<int>(2);                            // [0] targets, [1] scratch for the query sort
  g.q_cell_start = ar.take<int>((size_t)N * g.cells_stride);
  g.q_cell_of = ar.take<int>((size_t)N * n);
  g.q_order = ar.take<int>((size_t)N * n);
  g.total = ar.off;
  return g;
}

template <typename T>
int grid_build(const T* tgt, int N, int m, int S, const int32_t* m_valid, const GridWorkspace& g, cudaStream_t st) {
  const int cap = grid_cells_cap(m);
  DICP_CUDA(cudaMemsetAsync(g.cell_count, 0, sizeof(int) * (size_t)N * g.cells_stride, st));
  DICP_CUDA(cudaMemsetAsync(g.max_occupancy, 0, sizeof(int), st));
  const char* hs = getenv("DICP_GRID_HSCALE");            // tuning knob (tools/prof_grid.py); default GRID_HSCALE
  const double hscale = hs != nullptr ? atof(hs) : GRID_HSCALE;
  grid_bbox_kernel<T><<<N, 256, 0, st>>>(tgt, m, S, m_valid, cap, hscale, g.gparams, g.gdims);
  DICP_LAUNCH_CHECK();
  dim3 grid((m + 255) / 256, N);
  grid_count_kernel<T><<<grid, 256, 0, st>>>(tgt, m, S, m_valid, g.cells_stride, g.gparams, g.gdims, g.cell_count, g.cell_of);
  DICP_LAUNCH_CHECK();
  grid_scan_kernel<<<N, 1024, 0, st>>>(g.cell_count, g.cells_stride, g.gdims, g.cell_start, g.max_occupancy);
  DICP_LAUNCH_CHECK();
  DICP_CUDA(cudaMemsetAsync(g.cell_count, 0, sizeof(int) * (size_t)N * g.cells_stride, st));
  grid_scatter_kernel<T><<<grid, 256, 0, st>>>(tgt, m, S, m_valid, g.cells_stride, g.cell_of, g.cell_start, g.cell_count, (GridRec<T>*)g.recs);
  DICP_LAUNCH_CHECK();
  return 0;
}
template int grid_build<float>(const float*, int, int, int, const int32_t*, const GridWorkspace&, cudaStream_t);
template int grid_build<double>(const double*, int, int, int, const int32_t*, const GridWorkspace&, cudaStream_t);

// sort the queries by cell (once per call, at their initial position): g.q_order
template <typename T>
int grid_sort_queries(const T* q, const double* state, int N, int n, const GridWorkspace& g, cudaStream_t st) {
  DICP_CUDA(cudaMemsetAsync(g.cell_count, 0, sizeof(int) * (size_t)N * g.cells_stride, st));
  dim3 grid((n + 255) / 256, N);
  grid_qcount_kernel<T><<<grid, 256, 0, st>>>(q, state, n, g.cells_stride, g.gparams, g.gdims, g.cell_count, g.q_cell_of);
  DICP_LAUNCH_CHECK();
  grid_scan_kernel<<<N, 1024, 0, st>>>(g.cell_count, g.cells_stride, g.gdims, g.q_cell_start, g.max_occupancy + 1);
  DICP_LAUNCH_CHECK();
  DICP_CUDA(cudaMemsetAsync(g.cell_count, 0, sizeof(int) * (size_t)N * g.cells_stride, st));
  grid_qscatter_kernel<<<grid, 256, 0, st>>>(n, g.cells_stride, g.q_cell_of, g.q_cell_start, g.cell_count, g.q_order);
  DICP_LAUNCH_CHECK();
  return 0;
}
template int grid_sort_queries<float>(const float*, const double*, int, int, const GridWorkspace&, cudaStream_t);
template int grid_sort_queries<double>(const double*, const double*, int, int, const GridWorkspace&, cudaStream_t);

template <typename T>
int grid_search(const T* q, const double* state, int N, int n, int m, const GridWorkspace& g, bool sorted, int32_t* idx_out, T* d2_out,
                const int* stop_flag, cudaStream_t st, const T* tgt, int S, const int32_t* warm_idx) {
  dim3 grid((n + 127) / 128, N);
  if (tgt == nullptr) warm_idx = nullptr;
  grid_query_kernel<T><<<grid, 128, 0, st>>>(q, state, n, m, g.cells_stride, g.gparams, g.gdims, g.cell_start, (const GridRec<T>*)g.recs,
                                             sorted ? g.q_order : nullptr, idx_out, d2_out, stop_flag, tgt, S, warm_idx);
  DICP_LAUNCH_CHECK();
  return 0;
}
template int grid_search<float>(const float*, const double*, int, int, int, const GridWorkspace&, bool, int32_t*, float*, const int*, cudaStream_t, const float*, int, const int32_t*);
template int grid_search<double>(const double*, const double*, int, int, int, const GridWorkspace&, bool, int32_t*, double*, const int*, cudaStream_t, const double*, int, const int32_t*);

template <typename T>
static int nn_argmin_grid_impl(const T* q, const T* tgt, int N, int n, int m, int S, const int32_t* m_valid, int32_t* idx_out, T* d2_out,
                               void* ws, size_t ws_bytes, void* stream) {
  DICP_CHECK_ARG(q && tgt && idx_out, "null pointer");
  DICP_CHECK_ARG(N >= 1 && N <= 65535 && n >= 1 && m >= 1 && (S == 3 || S == 6), "bad sizes");
  const GridWorkspace g = grid_ws_layout(ws, ws_bytes, N, n, m, sizeof(T));
  if (ws == nullptr || g.total > ws_bytes) { set_error("dicp_nn_argmin_grid: workspace too small"); return DICP_ERR_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  int rc = grid_build<T>(tgt, N, m, S, m_valid, g, st);
  if (rc) return rc;
  rc = grid_sort_queries<T>(q, nullptr, N, n, g, st);
  if (rc) return rc;
  return grid_search<T>(q, nullptr, N, n, m, g, true, idx_out, d2_out, nullptr, st);
}

}  // namespace dicp

extern "C" {
size_t dicp_nn_grid_workspace_bytes(int N, int n, int m, int elem_size) { return dicp::grid_ws_layout(nullptr, 0, N, n, m, elem_size).total; }
int dicp_nn_argmin_grid_f32(const float* q, const float* tgt, int N, int n, int m, int S, const int32_t* m_valid, int32_t* idx_out,
                            float* d2_out, void* ws, size_t ws_bytes, void* stream) {
  return dicp::nn_argmin_grid_impl<float>(q, tgt, N, n, m, S, m_valid, idx_out, d2_out, ws, ws_bytes, stream);
}
int dicp_nn_argmin_grid_f64(const double* q, const double* tgt, int N, int n, int m, int S, const int32_t* m_valid, int32_t* idx_out,
                            double* d2_out, void* ws, size_t ws_bytes, void* stream) {
  return dicp::nn_argmin_grid_impl<double>(q, tgt, N, n, m, S, m_valid, idx_out, d2_out, ws, ws_bytes, stream);
}
}
