/*
 * dicp_b200.h — C ABI of the B200-native dICP hot path (libdicp_b200.so).
 *
 * The reference (utiasASRL/dICP) is pure Python/PyTorch and has no FFI of its own; its
 * boundary is the Python surface `dICP.ICP.ICP.icp` / `dICP.nn.nn.find_nn`.  The entry
 * points below are what a reference maintainer would bind (ctypes)
 * to replace the body of the loop at dICP/ICP.py:132-303 and the search at
 * dICP/nn.py:23-40.  Each declaration cites the reference lines it replaces.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *  - tensors are dense, row-major, in the element type of the suffix (_f32 / _f64);
 *  - functions allocate nothing: the caller passes a workspace sized by the matching
 *    `*_workspace_bytes` query (256-byte aligned);
 *  - all work is stream-ordered on `stream` (a cudaStream_t passed as void*);
 *  - return value: 0 ok, <0 argument error (DICP_ERR_*), >0 a cudaError_t;
 *    `dicp_last_error()` returns a thread-local message;
 *  - no C++ exceptions cross the boundary, no static mutable state, thread-safe.
 */
#ifndef DICP_B200_H
#define DICP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DICP_API __attribute__((visibility("default")))
#else
#define DICP_API
#endif

#define DICP_VERSION 201

#define DICP_ERR_ARG       -1   /* bad size / null pointer / unsupported option */
#define DICP_ERR_WORKSPACE -2   /* workspace too small */
#define DICP_ERR_HOOK      -4   /* conventional return value of a failing dicp_nn_hook */
#define DICP_ERR_SINGULAR  -3   /* a normal-equation matrix was exactly singular
                                   (the reference's torch.linalg.inv raises, ICP.py:201) */

#define DICP_PT2PL 0            /* icp_type 'pt2pl' (ICP.py:143-146, 172-176) */
#define DICP_PT2PT 1            /* icp_type 'pt2pt' (ICP.py:147-149, 177-180) */

#define DICP_NN_BRUTE 0         /* tiled brute-force scan (the default; nn.py:32-35 semantics)          */
#define DICP_NN_GRID  1         /* exact search through a per-pair uniform grid built once per call;
                                   same result as the scan, O(n) instead of O(n m) per iteration;
                                   falls back to the scan when a cloud is too clustered for a grid     */

#define DICP_LOSS_NONE   0      /* loss_fn=None            (ICP.py:157-158) */
#define DICP_LOSS_HUBER  1      /* loss.py:21-32 */
#define DICP_LOSS_CAUCHY 2      /* loss.py:34-41 */

/* doubles per (iteration, pair) record in the saved-state buffer, see DESIGN.md §3 */
#define DICP_STATE_STRIDE 64
/* slots of record [0][pair 0] that receive the call's status when it ends: the number of iterations run and
 * 1.0 if a normal-equation matrix was exactly singular (what a DICP_FLAG_NO_SYNC caller reads when it wants to) */
#define DICP_STATE_ITERS_RUN 62
#define DICP_STATE_SINGULAR  63

/* dicp_params.flags.  DICP_FLAG_NO_SYNC: with const_iter set, dicp_icp_forward_* only ENQUEUES the loop and returns
 * without synchronising the stream (the reference stops at every iteration, ICP.py:240, and torch.linalg.inv raises
 * on a singular matrix, ICP.py:201): *iters_run_host = max_iterations, a singular matrix is not reported by the
 * return value but left in saved_state[0][0][DICP_STATE_SINGULAR], and nn_mode GRID skips its host-side look at the
 * cell occupancy (a badly clustered cloud is then searched through the grid's bounded fallback instead of the scan).
 * Ignored when const_iter is 0. */
#define DICP_FLAG_NO_SYNC 1

/* Everything ICP.__init__ / ICP.dICP / config/dICP_config.yaml fix for one call. */
typedef struct dicp_params {
  int32_t icp_type;          /* DICP_PT2PL | DICP_PT2PT              ICP.py:30   */
  int32_t dim;               /* 2 | 3                                ICP.py:79   */
  int32_t differentiable;    /* soft (tanh / pseudo-Huber) weights   ICP.py:38   */
  int32_t loss_id;           /* DICP_LOSS_*                          ICP.py:158  */
  int32_t use_trim;          /* trim_dist is not None and >= 0       ICP.py:153  */
  int32_t max_iterations;    /*                                      ICP.py:31   */
  int32_t const_iter;        /* yaml parameters.const_iter           ICP.py:33   */
  int32_t check_every;       /* host polls the device "all converged" flag every this
                                many iterations when !const_iter (>=1; 1 == reference) */
  int32_t nn_mode;           /* DICP_NN_BRUTE | DICP_NN_GRID (not in the reference)         */
  int32_t flags;             /* DICP_FLAG_* bits, 0 == the reference's behaviour (not in the reference)     */
  double loss_metric;        /* loss_fn['metric']                    ICP.py:159  */
  double trim_dist;          /*                                      ICP.py:154  */
  double tanh_steepness;     /* yaml parameters.tanh_steepness       ICP.py:119  */
  double tolerance;          /*                                      ICP.py:32   */
  double match_ratio_thresh; /* yaml logging.matched_ratio_thresh    ICP.py:37   */
} dicp_params;

DICP_API int         dicp_version(void);
DICP_API const char* dicp_last_error(void);

/* ---- nearest neighbour: replaces cdist + argmin at dICP/nn.py:32-35 (== 83-86) --------
 * q   [N,n,3]   queries
 * tgt [N,m,S]   targets, S in {3,6}; only the first 3 columns are used
 * m_valid [N]   int32 or NULL: number of real rows of every target cloud; rows beyond it (the padding
 *               of a ragged batch) are never returned, whatever they contain
 * idx_out [N,n] int32 index of the nearest target (first minimum on exact ties)
 * d2_out  [N,n] squared distance to it, or NULL
 * The result is the argmin of the direct-form distance (x-y)^2 evaluated in the element type
 * (fp32: found through an expanded-form filter with a rigorous error margin, then refined /
 * rescanned exactly), so it equals the exact-arithmetic argmin whenever the relative gap
 * between the two smallest squared distances exceeds a few ulp.                          */
DICP_API size_t dicp_nn_workspace_bytes(int N, int n, int m, int elem_size);
DICP_API int dicp_nn_argmin_f32(const float* q, const float* tgt, int N, int n, int m, int S, const int32_t* m_valid,
                       int32_t* idx_out, float* d2_out, void* ws, size_t ws_bytes, void* stream);
DICP_API int dicp_nn_argmin_f64(const double* q, const double* tgt, int N, int n, int m, int S, const int32_t* m_valid,
                       int32_t* idx_out, double* d2_out, void* ws, size_t ws_bytes, void* stream);

/* Same result through a uniform grid: build (bounding box, counting sort)
 * + one-thread-per-query ring search.  ws sized by dicp_nn_grid_workspace_bytes.             */
DICP_API size_t dicp_nn_grid_workspace_bytes(int N, int n, int m, int elem_size);
DICP_API int dicp_nn_argmin_grid_f32(const float* q, const float* tgt, int N, int n, int m, int S, const int32_t* m_valid,
                            int32_t* idx_out, float* d2_out, void* ws, size_t ws_bytes, void* stream);
DICP_API int dicp_nn_argmin_grid_f64(const double* q, const double* tgt, int N, int n, int m, int S, const int32_t* m_valid,
                            int32_t* idx_out, double* d2_out, void* ws, size_t ws_bytes, void* stream);

/* The two stages of dicp_nn_argmin_*, exported for profiling and for callers that search the
 * same targets repeatedly (the solver loop packs once per call): pack the targets into
 * records padded with sentinels to a multiple of 32 rows (`dicp_nn_packed_bytes` bytes; fp32:
 * 16-byte filter records (-2y', |y'|^2) in the pair's translated frame y' = y - c, then the
 * 16-byte original (x,y,z,0) records, then the centres c; fp64: 32-byte (x,y,z,0)), then scan.
 * `state` (may be NULL): device records [N,DICP_STATE_STRIDE] doubles whose (C,r) transform the
 * queries first, x = C q + r, with the arithmetic of the solver loop (ICP.py:137).  d2_out (may
 * be NULL) receives the direct-form squared distance of the returned neighbour.  `ws` for the
 * scan is sized by dicp_nn_workspace_bytes.                                                  */
DICP_API size_t dicp_nn_packed_bytes(int N, int m, int elem_size);
DICP_API int dicp_nn_pack_f32(const float* tgt, int N, int m, int S, const int32_t* m_valid, void* packed, void* stream);
DICP_API int dicp_nn_pack_f64(const double* tgt, int N, int m, int S, const int32_t* m_valid, void* packed, void* stream);
DICP_API int dicp_nn_scan_f32(const float* q, const double* state, const void* packed, int N, int n, int m,
                     int32_t* idx_out, float* d2_out, void* ws, size_t ws_bytes, void* stream);
DICP_API int dicp_nn_scan_f64(const double* q, const double* state, const void* packed, int N, int n, int m,
                     int32_t* idx_out, double* d2_out, void* ws, size_t ws_bytes, void* stream);

/* Merge of slice results when ONE pair's targets are split over GPUs, over
 * NVLink peer memory instead of a collective library: every rank folds key = (bits(d2) << 32) |
 * (idx + offset) of each query into the key buffer of every rank (peer_bufs: HOST array of
 * `world` device addresses of count u64, peer-mapped, preset to all-ones) with a system-scope
 * 64-bit atomic MIN, starting at peer `first_peer`; reset_buf (may be NULL) is a local buffer
 * it also resets to all-ones for a later iteration.  After a cross-GPU barrier (the caller's),
 * dicp_nn_unpack_keys_f32 turns a rank's own buffer into indices (and distances).             */
DICP_API int dicp_nn_push_keys_f32(const float* d2, const int32_t* idx, int offset, size_t count,
                          const uint64_t* peer_bufs, int world, int first_peer, void* reset_buf, void* stream);
DICP_API int dicp_nn_unpack_keys_f32(const void* keys, size_t count, int32_t* idx_out, float* d2_out, void* stream);

/* FP32 CUDA-core probe for the roofline of the scan: launches a kernel of independent FFMA chains (8 CTAs of 256
 * threads per SM, 16 chains per thread, 8 x 16 FFMAs per loop iteration) and returns the FLOP count of the launch
 * (0 on error); the caller times it on `stream`.  sink: any device float.                                      */
DICP_API double dicp_probe_fp32_flops(int iters, float* sink, void* stream);

/* ---- neighbour gather and its adjoint: replaces repeat+gather at dICP/nn.py:37-38 ------
 * out[b,i,:] = tgt[b,idx[b,i],:]           grad_tgt[b,idx[b,i],:] += grad_out[b,i,:]
 * grad_tgt must be zero-initialised by the caller (it is accumulated into).             */
DICP_API int dicp_gather_rows_f32(const float* tgt, const int32_t* idx, int N, int n, int m, int S, float* out, void* stream);
DICP_API int dicp_gather_rows_f64(const double* tgt, const int32_t* idx, int N, int n, int m, int S, double* out, void* stream);
DICP_API int dicp_gather_rows_bwd_f32(const float* grad_out, const int32_t* idx, int N, int n, int m, int S, float* grad_tgt, void* stream);
DICP_API int dicp_gather_rows_bwd_f64(const double* grad_out, const int32_t* idx, int N, int n, int m, int S, double* grad_tgt, void* stream);

/* ---- the whole solver loop: replaces dICP/ICP.py:125-303 -------------------------------
 * inputs  source [N,n,3], target [N,m,S] (S=6 for pt2pl, 3 or 6 for pt2pt),
 *         T_init [N,4,4], w_init [N,n] (per point; pt2pt's x3 repeat is implicit)
 *         m_valid [N] int32 or NULL: number of real rows of every target cloud; rows beyond it
 *         are padding of a ragged batch (ICP.py:471-479) and are never returned as a neighbour,
 *         whatever they contain
 *         n_valid [N] int32 or NULL: number of real rows of every source cloud; rows beyond it are
 *         the zero padding of a ragged batch (ICP.py:361-426): they are never searched or evaluated
 *         (a block of padding costs nothing), their `weights` rows are 0, their `pc` rows the bare
 *         translation and nothing flows into them in the reverse pass -- what the reference obtains
 *         by giving zero points a zero weight.  Pass the same array to the backward call.
 * outputs pc [N,n,3], T [N,4,4],
 *         deltas  [N,K,6]        K = max_iterations rows are allocated; the first
 *         weights [N,K,n*R]      *iters_run_host are valid (R = 1 pt2pl, 3 pt2pt)
 *         costs   [N,K]
 *         converged [N] uint8, num_iters [N], match_ratio [N]      (ICP.py:288-292)
 * saved   saved_idx   [K,N,n] int32   correspondences of every iteration
 *         saved_state [(K+1),N,DICP_STATE_STRIDE] double  (C_k,r_k,active,delta_k,A_k ...)
 * iters_run_host: host int32, written before return (the call synchronises the stream
 *         once at its end, and every check_every iterations when !const_iter; not at all with
 *         const_iter and DICP_FLAG_NO_SYNC, see there).                                  */
DICP_API size_t dicp_icp_workspace_bytes(int N, int n, int m, int elem_size);
DICP_API int dicp_icp_forward_f32(const float* source, const float* target, const float* T_init, const float* w_init,
                         const int32_t* m_valid, const int32_t* n_valid, int N, int n, int m, int S, const dicp_params* params,
                         float* pc, float* T_out, float* deltas, float* weights, float* costs,
                         uint8_t* converged, float* num_iters, float* match_ratio,
                         int32_t* saved_idx, double* saved_state, int32_t* iters_run_host,
                         void* ws, size_t ws_bytes, void* stream);
DICP_API int dicp_icp_forward_f64(const double* source, const double* target, const double* T_init, const double* w_init,
                         const int32_t* m_valid, const int32_t* n_valid, int N, int n, int m, int S, const dicp_params* params,
                         double* pc, double* T_out, double* deltas, double* weights, double* costs,
                         uint8_t* converged, double* num_iters, double* match_ratio,
                         int32_t* saved_idx, double* saved_state, int32_t* iters_run_host,
                         void* ws, size_t ws_bytes, void* stream);

/* The same loop with the correspondence search supplied by the caller (second half:
 * one huge pair whose TARGETS are split over several GPUs).  Every iteration k the loop calls
 *     hook(user, k, state_k, idx_k)
 * on the host, in stream order: state_k is the device record [N,DICP_STATE_STRIDE] whose (C,r)
 * transform the source (pass it to dicp_nn_scan_*), idx_k the device [N,n] int32 the hook must
 * fill (by work enqueued on `stream`) before the accumulate kernel that follows reads it.  A
 * non-zero return aborts the call with that code.  params->nn_mode is ignored, nothing is
 * packed; everything else is identical to dicp_icp_forward_*.                              */
typedef int (*dicp_nn_hook)(void* user, int iteration, const double* state_k, int32_t* idx_k);
DICP_API int dicp_icp_forward_hooked_f32(const float* source, const float* target, const float* T_init, const float* w_init,
                         const int32_t* m_valid, const int32_t* n_valid, int N, int n, int m, int S, const dicp_params* params,
                         float* pc, float* T_out, float* deltas, float* weights, float* costs,
                         uint8_t* converged, float* num_iters, float* match_ratio,
                         int32_t* saved_idx, double* saved_state, int32_t* iters_run_host,
                         void* ws, size_t ws_bytes, void* stream, dicp_nn_hook hook, void* user);
DICP_API int dicp_icp_forward_hooked_f64(const double* source, const double* target, const double* T_init, const double* w_init,
                         const int32_t* m_valid, const int32_t* n_valid, int N, int n, int m, int S, const dicp_params* params,
                         double* pc, double* T_out, double* deltas, double* weights, double* costs,
                         uint8_t* converged, double* num_iters, double* match_ratio,
                         int32_t* saved_idx, double* saved_state, int32_t* iters_run_host,
                         void* ws, size_t ws_bytes, void* stream, dicp_nn_hook hook, void* user);

/* ---- reverse pass: replaces torch autograd through ICP.py:132-280 -----
 * grad_T [N,4,4], grad_pc [N,n,3] (either may be NULL == zero)
 * grad_source [N,n,3], grad_target [N,m,S], grad_w [N,n] (may be NULL), grad_T_init [N,4,4]
 * are overwritten.  iters_run is the value the forward returned.  The workspace is sized by
 * dicp_icp_backward_workspace_bytes (no search structures: a fraction of the forward's, which the
 * caller may free as soon as the forward call has returned).                              */
DICP_API size_t dicp_icp_backward_workspace_bytes(int N, int n, int m, int elem_size);
DICP_API int dicp_icp_backward_f32(const float* source, const float* target, const float* w_init, const int32_t* n_valid,
                          int N, int n, int m, int S, const dicp_params* params, int iters_run,
                          const int32_t* saved_idx, const double* saved_state,
                          const float* grad_T, const float* grad_pc,
                          float* grad_source, float* grad_target, float* grad_w, float* grad_T_init,
                          void* ws, size_t ws_bytes, void* stream);
DICP_API int dicp_icp_backward_f64(const double* source, const double* target, const double* w_init, const int32_t* n_valid,
                          int N, int n, int m, int S, const dicp_params* params, int iters_run,
                          const int32_t* saved_idx, const double* saved_state,
                          const double* grad_T, const double* grad_pc,
                          double* grad_source, double* grad_target, double* grad_w, double* grad_T_init,
                          void* ws, size_t ws_bytes, void* stream);

/* ---- single Gauss-Newton stage, exported for tests and profiling -------------------------
 * One iteration's accumulate (ICP.py:143-200) on given correspondences, exactly the kernel
 * launch of the loop.  partials_out [N, dicp_gn_blocks_per_pair(N, n), 32] doubles (31 used): one row per
 * persistent CTA that worked on the pair (unused rows are written as zeros), which the loop's update
 * kernel sums in index order; after the sum: upper-triangular A
 * (21 | 6) at 0, b (6 | 3) at 21, cost at 27, sum(w) 28, #(w>thresh) 29, #(w_init>thresh) 30.
 * C_r [N,12] = row-major C then r, in double.
 * rows (fp32 only, may be NULL): the targets re-laid by dicp_gn_pack_rows_f32 as one aligned
 * record per row (32 B (x,y,z,nx,ny,nz,0,0) for pt2pl, 16 B (x,y,z,0) for pt2pt; N*m*32 bytes
 * reserved = dicp_gn_rows_bytes) so that a correspondence is one 256-/128-bit load; the
 * loop builds this table once per call.  NULL gathers from `target` directly.               */
DICP_API int dicp_gn_blocks_per_pair(int N, int n);
DICP_API size_t dicp_gn_rows_bytes(int N, int m, int elem_size);
DICP_API int dicp_gn_pack_rows_f32(const float* target, int N, int m, int S, int icp_type, void* rows, void* stream);
DICP_API int dicp_gn_accumulate_f32(const float* source, const float* target, const void* rows, const float* w_init,
                           const int32_t* idx, const double* C_r, int N, int n, int m, int S,
                           const dicp_params* params, float* weights_out, double* partials_out, void* stream);
DICP_API int dicp_gn_accumulate_f64(const double* source, const double* target, const double* w_init, const int32_t* idx,
                           const double* C_r, int N, int n, int m, int S, const dicp_params* params,
                           double* weights_out, double* partials_out, void* stream);

/* The other half of one iteration (ICP.py:200-217, 239): sums the partial rows of
 * dicp_gn_accumulate_* in index order, solves A delta = -b (6x6, or 3x3 for dim 2, partial
 * pivoting in double, +1e-12 on the diagonal), applies C+ = exp(phi^)^T C, r+ = r - rho.
 * delta [N,6] in the compute type, C_r_next [N,12] doubles (row-major C+ then r+, rounded
 * through the compute type like the loop's state), below_tolerance [N] (may be NULL):
 * ||delta|| < params->tolerance.                                                          */
DICP_API int dicp_solve_update_f32(const double* partials, int nblk, const double* C_r, int N, const dicp_params* params,
                          float* delta, double* C_r_next, uint8_t* below_tolerance, void* stream);
DICP_API int dicp_solve_update_f64(const double* partials, int nblk, const double* C_r, int N, const dicp_params* params,
                          double* delta, double* C_r_next, uint8_t* below_tolerance, void* stream);

/* ---- Kabsch / SVD point-to-point step: the per-point half of dICP/ICP.py:557-563 ----------
 * sums[b] = ( sum_i p_i (3), sum_i q_i (3), sum_i q_i p_i^T (9, row-major), pad ),  q_i = tgt[b, idx[b,i], :3]
 * from which the centroids and the cross-covariance W = M/n - q_mean p_mean^T follow (stage API of the
 * loop below, exported for tests and profiling).  The _bwd entry is the adjoint used by autograd: grad_source is overwritten,
 * grad_target is zeroed and scatter-added into.                                              */
DICP_API size_t dicp_kabsch_workspace_bytes(int N, int n);
DICP_API int dicp_kabsch_sums_f32(const float* source, const float* target, const int32_t* idx, int N, int n, int m, int S,
                         float* sums /*[N,16]*/, void* ws, size_t ws_bytes, void* stream);
DICP_API int dicp_kabsch_sums_f64(const double* source, const double* target, const int32_t* idx, int N, int n, int m, int S,
                         double* sums /*[N,16]*/, void* ws, size_t ws_bytes, void* stream);
DICP_API int dicp_kabsch_sums_bwd_f32(const float* source, const float* target, const int32_t* idx, int N, int n, int m, int S,
                             const float* grad_sums, float* grad_source, float* grad_target, void* stream);
DICP_API int dicp_kabsch_sums_bwd_f64(const double* source, const double* target, const int32_t* idx, int N, int n, int m, int S,
                             const double* grad_sums, double* grad_source, double* grad_target, void* stream);

/* ---- the whole Kabsch / SVD point-to-point loop: replaces dICP/ICP.py:533-591 -------------------
 * Per iteration: search, fused sums, per-pair 3x3 SVD (one-sided Jacobi, double) + determinant fix +
 * composition, convergence test sum|C x + r - q|^2 < tolerance per pair (a batch stops when every pair
 * is below it; tolerance <= 0: no test, max_iterations are run).  literal != 0: C = U D V as ICP.py:570
 * writes it, else the Kabsch rotation U D V^T.  T_init is only composed into T_out (ICP.py:545-547).
 * saved_idx [K,N,n] int32, saved_state [(K+1),N,DICP_STATE_STRIDE] doubles; *iters_run_host is written
 * before return (one stream synchronisation at the end, plus one every check_every iterations when a
 * tolerance is set).  The _backward entry is the hand-written reverse pass: grad_T / grad_pc may be NULL,
 * grad_source, grad_target, grad_T_init are overwritten.                                              */
DICP_API size_t dicp_kabsch_icp_workspace_bytes(int N, int n, int m, int elem_size);
DICP_API int dicp_kabsch_icp_forward_f32(const float* source, const float* target, const float* T_init, int N, int n, int m, int S,
                                int max_iterations, double tolerance, int literal, int check_every, float* pc, float* T_out,
                                int32_t* saved_idx, double* saved_state, int32_t* iters_run_host, void* ws, size_t ws_bytes, void* stream);
DICP_API int dicp_kabsch_icp_forward_f64(const double* source, const double* target, const double* T_init, int N, int n, int m, int S,
                                int max_iterations, double tolerance, int literal, int check_every, double* pc, double* T_out,
                                int32_t* saved_idx, double* saved_state, int32_t* iters_run_host, void* ws, size_t ws_bytes, void* stream);
DICP_API int dicp_kabsch_icp_backward_f32(const float* source, const float* target, const float* T_init, int N, int n, int m, int S,
                                 int literal, int iters_run, const int32_t* saved_idx, const double* saved_state,
                                 const float* grad_T, const float* grad_pc, float* grad_source, float* grad_target,
                                 float* grad_T_init, void* ws, size_t ws_bytes, void* stream);
DICP_API int dicp_kabsch_icp_backward_f64(const double* source, const double* target, const double* T_init, int N, int n, int m, int S,
                                 int literal, int iters_run, const int32_t* saved_idx, const double* saved_state,
                                 const double* grad_T, const double* grad_pc, double* grad_source, double* grad_target,
                                 double* grad_T_init, void* ws, size_t ws_bytes, void* stream);

/* ---- Gumbel-softmax soft nearest neighbour: replaces dICP/nn.py:43-70 ----------------------
 * out[b,i,:] = sum_j softmax_j((-|x_i - y_j|^2 + G_ij) / tau) y[b,j,:],  G = -log(-log(U+eps)+eps)
 * streamed with an online softmax (no (N,n,m) tensor); U comes from Philox4x32-10 keyed by `seed`
 * with counter (pair, query, target/4), so a call is reproducible and the backward regenerates
 * the same noise.  lse [N,n] (log-sum-exp of the logits) is saved for the backward; ws_n is an
 * [N,n] scratch array.  dicp_gumbel_noise_* writes the noise tensor itself (tests only).         */
DICP_API int dicp_gumbel_nn_f32(const float* x, const float* y, int N, int n, int m, int S, unsigned long long seed,
                       double eps, double tau, float* out, float* lse, void* stream);
DICP_API int dicp_gumbel_nn_f64(const double* x, const double* y, int N, int n, int m, int S, unsigned long long seed,
                       double eps, double tau, double* out, double* lse, void* stream);
DICP_API int dicp_gumbel_nn_bwd_f32(const float* x, const float* y, const float* out, const float* lse, const float* grad_out,
                           int N, int n, int m, int S, unsigned long long seed, double eps, double tau,
                           float* grad_x, float* grad_y, float* ws_n, void* stream);
DICP_API int dicp_gumbel_nn_bwd_f64(const double* x, const double* y, const double* out, const double* lse, const double* grad_out,
                           int N, int n, int m, int S, unsigned long long seed, double eps, double tau,
                           double* grad_x, double* grad_y, double* ws_n, void* stream);
DICP_API int dicp_gumbel_noise_f32(int N, int n, int m, unsigned long long seed, double eps, float* G, void* stream);
DICP_API int dicp_gumbel_noise_f64(int N, int n, int m, unsigned long long seed, double eps, double* G, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DICP_B200_H */
