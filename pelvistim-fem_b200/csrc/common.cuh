// common.cuh — shared host/device plumbing of libptfem (contexts, buffers, error handling).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/ptfem.h"
#include "window.cuh"

namespace ptfem {

extern thread_local std::string g_err;

int set_err(int code, const char* fmt, ...);

#define PT_CK(expr)                                                                              \
  do {                                                                                           \
    cudaError_t e__ = (expr);                                                                    \
    if (e__ != cudaSuccess)                                                                      \
      return ptfem::set_err(PTFEM_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,        \
                            cudaGetErrorString(e__));                                            \
  } while (0)

#define PT_TRY(expr)              \
  do {                            \
    int r__ = (expr);             \
    if (r__ != PTFEM_OK) return r__; \
  } while (0)

#define PT_ARG(cond, msg)                                               \
  do {                                                                  \
    if (!(cond)) return ptfem::set_err(PTFEM_ERR_ARG, "%s: %s", __func__, msg); \
  } while (0)

// Device allocations go through a small caching allocator (util.cu): cudaFree synchronises the device and
// costs ~0.3 s per mesh of the size-L working set, which an end-to-end sweep step would pay every time.
// Freed blocks are kept and handed out again (best fit within 25 %); the cache is flushed when cudaMalloc
// runs out of memory and when the last context is destroyed.
int dev_alloc(void** p, size_t bytes);
void dev_free(void* p);
void dev_cache_flush();

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  int alloc(size_t count) {
    if (count <= n && p) return PTFEM_OK;
    release();
    if (count == 0) count = 1;
    const int rc = dev_alloc((void**)&p, count * sizeof(T));
    if (rc != PTFEM_OK) {
      p = nullptr;
      return rc;
    }
    n = count;
    return PTFEM_OK;
  }
  void release() {
    if (p) dev_free(p);
    p = nullptr;
    n = 0;
  }
  ~DevBuf() { release(); }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
};

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace ptfem

// Supported padded system counts (vectors are stored interleaved [nn][S]).
static inline int ptfem_pad_nsys(int s) {
  if (s <= 1) return 1;
  if (s <= 2) return 2;
  if (s <= 4) return 4;
  if (s <= 8) return 8;
  if (s <= 16) return 16;
  return -1;
}

struct NcclApi;           // dist.cu
struct CoarseSpace;       // coarse.cuh
struct ptfem_dist_state;  // dist.cu

struct ptfem_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;  // copies / halo
  cudaStream_t stream3 = nullptr;  // mesh uploads (host -> device), so they overlap both the kernels and the read-backs
  cudaStream_t stream_x = nullptr; // coarse-grid PCG: x += alpha p beside the grid hierarchy (PTFEM_SPLIT_X), forked and joined by events
  cudaEvent_t ev_xfork = nullptr, ev_xjoin = nullptr;
  int64_t launches = 0;
  cudaEvent_t ev_phi_ready = nullptr;
  cudaEvent_t ev_j_ready = nullptr, ev_j_copied = nullptr;  // asynchronous read-back of the nodal current (stream2)
  double* h_pinned = nullptr;  // small pinned scratch (scalars)
  size_t h_pinned_n = 0;
  int tune_interleave = 1;         // SpMV work distribution: 1 moving front, 0 contiguous run per CTA (PTFEM_INTERLEAVE)
  int tune_stream_cap = 0;         // staged entries per tile (PTFEM_STREAM_CAP, 0 = default)
  int tune_stream_rows = 0;        // most rows per tile (PTFEM_STREAM_ROWS, 0 = 128)
  int tune_stream_tpr = 1;         // lanes per row of the streaming SpMV (PTFEM_STREAM_TPR: 1, 2, 4, 8)
  int tune_stream_stages = 2;      // shared-memory stages of the streaming SpMV (PTFEM_STREAM_STAGES: 2, 3)
  int tune_morton = -1;            // streaming SpMV walks rows along a Morton curve: -1 auto, 0 off, 1 on (PTFEM_MORTON)
  int tune_p2p_fused = 0;          // row-partitioned solve: SpMV loads halo entries from peer memory itself (PTFEM_P2P_FUSED)
  int tune_xprefetch = 0;          // streaming SpMV prefetches the leading edge of x into L2 (PTFEM_XPREFETCH)
  int tune_restrict_occ = 4;       // PTFEM_RESTRICT_OCC: resident CTAs per SM the restriction kernel is compiled for (2, 4, 6; 4 measured best in round 1)
  int tune_coarse_fused = 0;       // PTFEM_COARSE_FUSED: grid hierarchy of the coarse-grid preconditioner as one cooperative kernel
  double tune_coarse_weight = 0.0; // PTFEM_COARSE_WEIGHT: weight of the coarse-grid levels against the Jacobi term (0 = 2 / (levels + 1))
  int tune_pair = 0;               // PTFEM_SPMM_PAIR: multi-RHS streaming SpMM reads two non-zeros per shared-memory access (even-padded copy); measured slower (0.335 vs 0.289 ms): off
  int tune_chain_tail = 0;         // PTFEM_CHAIN_TAIL: partitioned solve runs the replicated smallest grid levels as one block (measured slower: off)
  int tune_pupdate_occ = 4;        // PTFEM_PUPDATE_OCC: resident CTAs per SM the one-pair p-update is compiled for (4, 5, 6)
  int tune_pupdate_grid = 0;       // PTFEM_PUPDATE_GRID: CTAs per SM of the coarse-grid p-update's grid (0 = 8, two waves of the 4 resident)
  int tune_pupdate_np = 1;         // PTFEM_PUPDATE_NP: pairs per trip of the coarse-grid p-update (1: 4 CTAs/SM, 2: 2 CTAs/SM, all loads of both first)
  int tune_fuse_update = 1;        // PTFEM_FUSE_UPDATE: CG residual update inside the restriction's gather (coarse-grid PCG, one matrix)
  int tune_fuse_occ = 3;           // PTFEM_FUSE_OCC: resident CTAs per SM the fused update + restriction is compiled for (3: 80 registers, 4: 64)
  int tune_fuse_grid = 0;          // PTFEM_FUSE_GRID: its CTAs per SM (0 = one wave of the resident CTAs; at most 8: the CG partial sums hold 8 per SM)
  int tune_restrict_grid = 0;      // PTFEM_RESTRICT_GRID: CTAs per SM of the plain restriction's grid (0 = 32: eight waves of the 4 resident)
  int tune_fuse_pipe = 0;          // PTFEM_FUSE_PIPE: the fused update + restriction as a cp.async software pipeline (contiguous task runs per warp, next trip's rows in flight)
  int tune_fuse_prefetch = 0;      // PTFEM_FUSE_PREFETCH: the fused update + restriction asks the next trip's rows of r, q, dinv into L2 when their indices arrive
  int tune_split_x = 0;            // PTFEM_SPLIT_X: coarse-grid PCG updates x on a side stream while the grid hierarchy runs (1: forked after the product, 2: after the restriction); the p-update then leaves x alone
  int tune_split_x_ctas = 0;       // PTFEM_SPLIT_X_CTAS: resident CTAs per SM of the side-stream x-update (0 = 2)
  int tune_window = 1;             // PTFEM_SPMM_WINDOW: multi-RHS SpMM out of shared-memory x windows where the numbering allows a plan (window.cu); 0 = streaming kernel
  int tune_window_bx = 0;          // PTFEM_WINDOW_BX: rows of a brick along a line (0 = 16)
  int tune_window_ctas = 0;        // PTFEM_WINDOW_CTAS: resident CTAs per SM of the window SpMM (0 = as many as fit, at most 4)
  int tune_ctas_per_sm = 0;        // cap on resident CTAs per SM of the streaming SpMV (PTFEM_CTAS_PER_SM)
  std::unordered_map<const void*, size_t> func_smem;  // dynamic shared memory limit raised per kernel
  // NCCL (row-partitioned solves)
  NcclApi* nccl = nullptr;
  void* comm = nullptr;
  int rank = 0, nranks = 1;
};

// One CG workspace per (mesh, S).
struct PcgWork {
  int S = 0;  // padded systems
  ptfem::DevBuf<double> r, p, q;       // CG vectors [nn][S]
  ptfem::DevBuf<double> z, rt, d;      // Chebyshev preconditioner vectors
  ptfem::DevBuf<double> coef;          // Chebyshev coefficients [(deg+1)][S][2]
  ptfem::DevBuf<double> partial;       // per-CTA partial sums
  ptfem::DevBuf<double> scal;          // device scalars, see solver.cu (SC_*), each [16]
  ptfem::DevBuf<unsigned int> ticket;  // last-CTA-done counter
  cudaGraphExec_t graph = nullptr;     // check_every captured iterations
  int graph_iters = 0;
  int graph_variant = -1;
  int graph_precond = -1;
  int graph_cheb = 0;
  int64_t graph_coarse = -1;
  int64_t graph_launches = 0;
  const void* graph_x = nullptr;
  const void* graph_val = nullptr;
  const void* graph_b = nullptr;
  const void* graph_dinv = nullptr;
};

struct ptfem_mesh {
  ptfem_ctx* ctx = nullptr;
  bool upload_pending = false;    // ptfem_mesh_create_async: arrays still arriving on stream3, validation and bounding box still to do
  cudaEvent_t ev_upload = nullptr;
  int64_t nn = 0, nt = 0, nb = 0, nnz = 0;
  // mesh
  ptfem::DevBuf<double> xyz;     // [nn][3]
  ptfem::DevBuf<int32_t> tets;   // [nt][4]
  ptfem::DevBuf<int32_t> region; // [nt]
  ptfem::DevBuf<int32_t> tris;   // [nb][3]
  ptfem::DevBuf<int32_t> bcid;   // [nb]
  // pattern
  bool has_pattern = false;
  ptfem::DevBuf<int32_t> n2t_ptr, n2t;   // node -> tets (sorted)
  ptfem::DevBuf<int32_t> n2b_ptr, n2b;   // node -> boundary tris (sorted)
  ptfem::DevBuf<int32_t> rowptr, col, diag;
  ptfem::DevBuf<int32_t> e2nnz;          // [nt][16]
  ptfem::DevBuf<int32_t> gptr, gsrc;     // nnz -> (tet*16 + ij) contributions, sorted
  // stream-SpMV row blocks
  // streaming SpMV: Morton processing order (rowid[j] = mesh row handled j-th) and the matrix copy in that order
  bool has_rowperm = false;
  double row_coherence = 1.0;     // fraction of rows whose columns are the previous row's shifted by one
  ptfem::DevBuf<int32_t> rowid, prowptr, pcol;
  ptfem::DevBuf<double> pval;     // [nnz] values of val_bc in processing order (single matrix only)
  // multi-RHS streaming SpMM: copy of the matrix with every row padded to an EVEN length (pad: value 0, column = the row's
  // own), so that a lane takes two non-zeros per shared-memory access (one 16-byte load of values, one 8-byte load of columns)
  bool has_qcopy = false;
  int64_t qnnz = 0;
  ptfem::DevBuf<int32_t> qrowptr, qcol;
  ptfem::DevBuf<double> qval;
  int32_t q_rows = 0, q_cap = 0;  // its tile geometry
  // window SpMM (window.cuh): plan + the buffers it points into
  WindowPlan win;
  ptfem::DevBuf<WinTile> win_tiles;
  ptfem::DevBuf<WinRange> win_ranges;
  ptfem::DevBuf<unsigned char> win_blob;
  ptfem::DevBuf<int32_t> win_rowid;   // brick order: processing position -> mesh row
  ptfem::DevBuf<int32_t> win_rowtile, win_trow0;   // tile of a processing row; first processing row of a tile
  double bb_lo[3] = {0, 0, 0}, bb_hi[3] = {0, 0, 0};
  int32_t stream_rows = 0;        // rows per tile of the streaming SpMV (0 = not usable on this pattern)
  int32_t stream_cap = 0;         // staged entries per tile
  int32_t max_row = 0;            // longest row of the pattern
  double h_max = 0.0;             // longest tet edge
  // geometry factors
  bool has_geom = false;
  ptfem::DevBuf<double> G;        // [nt][10]  |V| gradNi.gradNj (i<=j)
  ptfem::DevBuf<double> vol;      // [nt]      |V|
  ptfem::DevBuf<double> tri_area; // [nb]
  ptfem::DevBuf<double> mlump;    // [nn] lumped mass (sum V/4)
  ptfem::DevBuf<int32_t> valence; // [nn] tets touching the node
  // assembled systems
  int nval = 0;                   // value sets assembled (1 or nsys), padded: nvalp
  int nvalp = 0;
  int nreg = 0;
  ptfem::DevBuf<uint8_t> regidx;  // [nt] dense region index
  ptfem::DevBuf<double> sigma_tab;// [nvalp][nreg]
  ptfem::DevBuf<double> val_raw;  // [nnz][nvalp]
  ptfem::DevBuf<double> val_bc;   // [nnz][nvalp]
  // boundary conditions
  int nrhs = 0, nrhsp = 0;        // right-hand sides (padded)
  ptfem::DevBuf<uint8_t> isdir;   // [nn]
  ptfem::DevBuf<double> dirval;   // [nn][nrhsp]
  ptfem::DevBuf<double> tri_load; // [nrhsp][nb]
  ptfem::DevBuf<double> b_neu;    // [nn][nrhsp]  Neumann load vector
  ptfem::DevBuf<double> b;        // [nn][S]      rhs after elimination
  ptfem::DevBuf<double> dinv;     // [nn][nvalp]  inverse diagonal of the eliminated matrix
  bool bc_dirty = true;
  int64_t matrix_epoch = 0;       // bumped whenever val_bc is rewritten (coarse Galerkin operators follow it)
  CoarseSpace* coarse = nullptr;  // geometric coarse spaces of the two-level preconditioner (coarse.cu)
  // solution / fields
  int S = 0;                      // systems of the last solve (padded)
  int nsys_user = 0;
  ptfem::DevBuf<double> phi;      // [nn][S]
  ptfem::DevBuf<double> Jnode;    // [nn][3] last recovered nodal current
  ptfem::DevBuf<double> Eelem, Jelem; // [nt][3]
  ptfem::DevBuf<double> mval;     // [nnz] consistent mass values (L2 recovery)
  ptfem::DevBuf<double> mdinv, mrhs, mx; // mass-matrix Jacobi, rhs [nn][4], solution [nn][4]
  ptfem::DevBuf<double> phis;     // [nn] VTK-smoothed potential (ROI metric)
  int J_sys = -1;                 // system whose nodal current is current (in Jnode, or a block of Jall when J_from_all)
  ptfem::DevBuf<double> Jall;     // [nsys][nn][3] nodal currents of every system (ptfem_recover_current_batch)
  ptfem::DevBuf<double> Jelem_all;// [nt][3][S] element currents of every system (scratch of the batched recovery)
  bool J_all_valid = false;       // Jall holds the currents of the solution on the device
  int J_all_method = -1;
  bool J_from_all = false;        // the current system's nodal current is Jall + J_sys*nn*3
  ptfem::DevBuf<float> tcen;      // [nt][4] tet centroids in single precision (prefilter of the ROI scans)
  ptfem::DevBuf<double> nchunk;   // [ceil(nn/128)][4] bounding sphere of every run of 128 consecutive nodes (ROI smoothing)
  ptfem::DevBuf<float> tchunk;    // [ceil(nt/256)][4] bounding sphere (centre, radius) of every run of 256 consecutive centroids
  bool has_tcen = false;
  ptfem::DevBuf<double> phis_all; // [nroi][nn] smoothed potentials of a metric batch
  bool j_copy_pending = false;    // an asynchronous device->host copy of Jnode may still be reading it
  int mass_iters = 0;
  ptfem::DevBuf<double> scratch_phi;  // [nsys][nn] staging of ptfem_phi_get_all_async (scratch_d is reused by the metric kernels)
  ptfem::DevBuf<double> scratch_d;
  ptfem::DevBuf<int32_t> scratch_i;
  PcgWork work;
  PcgWork work3;                  // mass-matrix solves (S=4)
  // distributed system (row block)
  bool is_part = false;           // one rank's part of a larger mesh (ptfem_mesh_set_bbox): may hold no Dirichlet node
  bool is_dist = false;
  int64_t nloc = 0, nhalo = 0;
  int nnbr = 0;
  std::vector<int32_t> nbr_rank, send_ptr, recv_ptr;
  ptfem::DevBuf<int32_t> send_idx;
  ptfem::DevBuf<double> send_buf;
  ptfem_dist_state* dist = nullptr;
};

// ---- helpers implemented in util.cu ---------------------------------------------------------
namespace ptfem {
int exclusive_scan_i32(ptfem_ctx* ctx, const int32_t* in, int32_t* out, int64_t n, int64_t* total);
int fill_i32(ptfem_ctx* ctx, int32_t* p, int32_t v, int64_t n);
int fill_f64(ptfem_ctx* ctx, double* p, double v, int64_t n);
// window.cu: plan of the window SpMM for this pattern (m->win.valid says whether there is one), and its values
int window_plan_build(ptfem_mesh* m);
int window_refresh_values(ptfem_mesh* m, const double* val);
}  // namespace ptfem

#define PT_LAUNCH_CHECK(ctx)                                                                     \
  do {                                                                                           \
    (ctx)->launches++;                                                                           \
    cudaError_t e__ = cudaGetLastError();                                                        \
    if (e__ != cudaSuccess)                                                                      \
      return ptfem::set_err(PTFEM_ERR_CUDA, "%s:%d: kernel launch -> %s", __FILE__, __LINE__,    \
                            cudaGetErrorString(e__));                                            \
  } while (0)

// ---- device helpers -----------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
