// coarse.cuh — geometric coarse spaces of the two-level / multilevel-additive PCG preconditioner (coarse.cu).
//
//   M^-1 r = D^-1 r + sum_l Z_l B_l Z_l^T r
//
// Z_l interpolates trilinearly from a regular grid laid over the mesh bounding box to the mesh nodes (rows of
// Dirichlet nodes are zero).  The coarsest grid carries the exact Galerkin inverse B = (Z^T K Z)^-1 (dense, a few
// thousand unknowns), the finer ones only the inverse diagonal of their Galerkin operator (BPX).  The reference
// solves with a direct method (UMFPACK, step01_box/case.sif:41-42), so any SPD preconditioner gives the same
// answer; this one cuts the Jacobi-PCG iteration count of the refined layered slab by about an order of magnitude.
#pragma once
#include "common.cuh"

constexpr int kMaxCoarseLevels = 6;
constexpr int kDefaultCoarseNodes = 300;   // unknowns of the exactly inverted grid (see coarse_prepare)

struct CoarseGrid {  // passed by value to kernels
  int n[3];          // cells per axis
  double lo[3], inv_h[3];
};

// what the CG kernels need to add the coarse correction to z (by value)
struct CoarseDev {
  int nx1 = 1, ny1 = 1;         // grid nodes per axis (x, y) of the finest level
  int shift = 0;                // its cells are 2^shift smaller than the table's
  const double* y = nullptr;    // [k_0][S] sum over the levels, on the finest grid
  const float4* ctab = nullptr; // [nn] see CoarseSpace::ctab
};

struct CoarseLevel {
  CoarseGrid g;
  int64_t k = 0, ncell = 0;   // grid nodes, cells
  bool exact = false;
  int kp = 0;                 // k padded to the tile of the dense inverse (exact level)
  ptfem::DevBuf<int32_t> rows, cellptr;   // mesh rows sorted by cell, [ncell+1]
  ptfem::DevBuf<double> binv;             // exact: [kp][kp] inverse; else [k] inverse diagonal
  int split = 1;                          // CTAs sharing one cell in the restriction
  int shift = 0;                          // cells are 2^shift smaller than the coarsest level's
  ptfem::DevBuf<double> part;             // [ncell][split][8][S] restriction partials
  ptfem::DevBuf<double> rc, yc;           // [kp or k][S]
  ptfem::DevBuf<double> yt;               // [k][S] y of this level + interpolated coarser levels
};

struct CoarseSpace {
  int nlev = 0;
  CoarseLevel lev[kMaxCoarseLevels];
  bool geom_ok = false;
  int64_t matrix_epoch = -1;  // ptfem_mesh::matrix_epoch the Galerkin operators were built for
  int64_t generation = 0;     // bumped whenever buffers or grids change (CUDA-graph key)
  int S = 0;
  int VS = 1;                 // Galerkin operators per level: 1 (one matrix) or S (batched matrices, stored one after the other)
  int req_nodes = 0, req_levels = 0;
  ptfem::DevBuf<double> cdot;             // [nlev][16] r_c . y_c per level and system
  ptfem::DevBuf<double> dpart;            // per-CTA partials of those dots
  ptfem::DevBuf<unsigned int> ticket;
  // per mesh row, 16 bytes: position in the coarsest grid {t_x, t_y, t_z as float, packed cell (10 bits per axis) with
  // bit 31 set for a Dirichlet row}; finer levels derive theirs by doubling.  Single precision halves what restriction and
  // prolongation read per row; every consumer (Galerkin operators, restriction, prolongation) decodes the SAME rounded
  // coordinates, so Z is one matrix everywhere and M^-1 stays symmetric.  Rebuilt with the matrix (Dirichlet flags).
  ptfem::DevBuf<float4> ctab;
  ptfem::DevBuf<float4> ctab0;            // the same rows in the order of level 0's row list (restriction reads it in step with the list)
  ptfem::DevBuf<int32_t> flag;            // [0] non-positive pivot seen, [1] slow-path entries
  double setup_ms = 0.0;
  // distributed set-up (a rank's local mesh = owned rows first, then ghost nodes): only the first row_limit rows own
  // Galerkin contributions and enter the row lists; nn_levels = node count of the WHOLE mesh (level rule); partial_mode:
  // coarse_prepare stops at the raw Galerkin sums, which the ranks add up before coarse_finish_sums inverts them
  int64_t row_limit = -1, nn_levels = -1;
  bool partial_mode = false, partial_ready = false;
  int chain_grid = 0;                     // CTAs of the cooperative grid-hierarchy kernel (0: separate kernels)
  ptfem::DevBuf<double> chain_part;       // its per-CTA dot partials [nlev][chain_grid][S]
};

__device__ __forceinline__ void coarse_locate(const CoarseGrid& g, const double* __restrict__ xyz, int64_t i, int (&c)[3],
                                              double (&t)[3]) {
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double u = (__ldg(xyz + 3 * i + d) - g.lo[d]) * g.inv_h[d];
    int ci = (int)floor(u);
    ci = ci < 0 ? 0 : (ci > g.n[d] - 1 ? g.n[d] - 1 : ci);
    double tt = u - (double)ci;
    // snap onto the grid planes: a mesh node on a plane must give its neighbour across the plane weight 0, not 1e-12
    // (a coarse function carried only by such weights would make the Galerkin matrix numerically singular)
    t[d] = tt < 1e-9 ? 0.0 : (tt > 1.0 - 1e-9 ? 1.0 : tt);
    c[d] = ci;
  }
}
// table row -> cell and local coordinates on the level whose cells are 2^shift smaller; false for Dirichlet rows.
// Split in load + decode so that callers can issue the loads of several rows before any of them is used; the decode
// is branch-free (a Dirichlet row decodes to cell 0 and live = false).
struct CoarseRaw {
  float4 v;
};
constexpr unsigned int kCoarseDirBit = 0x80000000u;
constexpr int kCoarseMaxCells = 1023;   // cells per axis of the coarsest grid that fit the packed entry
__device__ __forceinline__ float4 coarse_row_pack(const int (&c)[3], const double (&t)[3], bool dirichlet) {
  const unsigned int bits = (unsigned)c[0] | ((unsigned)c[1] << 10) | ((unsigned)c[2] << 20) | (dirichlet ? kCoarseDirBit : 0u);
  return make_float4((float)t[0], (float)t[1], (float)t[2], __uint_as_float(bits));
}
__device__ __forceinline__ CoarseRaw coarse_row_load(const float4* __restrict__ ctab, int64_t i) {
  CoarseRaw r;
  r.v = __ldg(ctab + i);
  return r;
}
__device__ __forceinline__ bool coarse_row_decode(const CoarseRaw& r, int shift, int (&c)[3], double (&t)[3]) {
  unsigned int cell = __float_as_uint(r.v.w);
  const bool live = !(cell & kCoarseDirBit);
  cell = live ? cell : 0u;
  c[0] = (int)(cell & 0x3ffu);
  c[1] = (int)((cell >> 10) & 0x3ffu);
  c[2] = (int)((cell >> 20) & 0x3ffu);
  t[0] = (double)r.v.x;
  t[1] = (double)r.v.y;
  t[2] = (double)r.v.z;
  if (shift > 0) {
    const int f = 1 << shift;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const double u = t[d] * (double)f;
      int k = (int)u;
      k = k > f - 1 ? f - 1 : k;
      t[d] = u - (double)k;
      c[d] = c[d] * f + k;
    }
  }
  return live;
}
__device__ __forceinline__ bool coarse_row(const float4* __restrict__ ctab, int64_t i, int shift, int (&c)[3], double (&t)[3]) {
  return coarse_row_decode(coarse_row_load(ctab, i), shift, c, t);
}
// the eight trilinear weights from three subtractions and twelve products
__device__ __forceinline__ void coarse_weights(const double (&t)[3], double (&w)[8]) {
  const double x0 = 1.0 - t[0], y0 = 1.0 - t[1], z0 = 1.0 - t[2];
  const double xy[4] = {x0 * y0, t[0] * y0, x0 * t[1], t[0] * t[1]};
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    w[a] = xy[a] * z0;
    w[a + 4] = xy[a] * t[2];
  }
}
__device__ __forceinline__ double coarse_weight(const double (&t)[3], int a) {
  return ((a & 1) ? t[0] : 1.0 - t[0]) * ((a & 2) ? t[1] : 1.0 - t[1]) * ((a & 4) ? t[2] : 1.0 - t[2]);
}
__device__ __forceinline__ int64_t coarse_node(const CoarseGrid& g, const int (&c)[3], int a) {
  return (int64_t)(c[0] + (a & 1)) + (int64_t)(g.n[0] + 1) * ((c[1] + ((a >> 1) & 1)) + (int64_t)(g.n[1] + 1) * (c[2] + (a >> 2)));
}

// (Z y)[row], systems s .. s+NS-1 (y = sum of all levels on the finest grid); raw = the row's table entry
template <int S, int NS>
__device__ __forceinline__ void coarse_prolong(const CoarseDev& cd, const CoarseRaw& raw, int s, double (&out)[NS]) {
  int c[3];
  double t[3], w[8];
  const double live = coarse_row_decode(raw, cd.shift, c, t) ? 1.0 : 0.0;
  coarse_weights(t, w);
  const int64_t n0 = c[0] + (int64_t)cd.nx1 * (c[1] + (int64_t)cd.ny1 * c[2]);
  double acc[NS];
#pragma unroll
  for (int k = 0; k < NS; ++k) acc[k] = 0.0;
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int64_t node = n0 + (a & 1) + (int64_t)cd.nx1 * (((a >> 1) & 1) + (int64_t)cd.ny1 * (a >> 2));
    const double* y = cd.y + (size_t)node * S + s;
    if constexpr (NS == 2) {
      const double2 v = __ldg(reinterpret_cast<const double2*>(y));
      acc[0] = fma(w[a], v.x, acc[0]);
      acc[1] = fma(w[a], v.y, acc[1]);
    } else {
      acc[0] = fma(w[a], __ldg(y), acc[0]);
    }
  }
#pragma unroll
  for (int k = 0; k < NS; ++k) out[k] = acc[k] * live;
}

// "this rank's exchange buffer is complete" signal of the sharded coarse exchange, raised by the last CTA of the kernel that
// completes the buffer (dist.cu fills it in; n == 0: no signal)
struct CoarseSignal {
  unsigned long long* seq = nullptr;          // local sequence counter (Mail::xred), bumped by one
  unsigned long long* peer_flag[16] = {};     // xready_from[this rank] in every other rank's mailbox
  int n = 0;
  unsigned int* ticket = nullptr;
};

namespace ptfem {
// (re)builds what is stale: grids + sorted row lists (mesh coordinates / requested size changed) and the Galerkin
// operators (matrix changed).  target_nodes: unknowns of the exact level (0 = default); extra_levels: finer
// diagonal-only levels (each halves the cell size).
int coarse_prepare(ptfem_mesh* m, int target_nodes, int extra_levels, int S);
// distributed set-up: raw Galerkin sums of the owned rows [kp*kp | k_0 | k_1 | ...] (size query with out == nullptr), and the
// inversion once the launcher has summed them over the ranks
int coarse_partial_sums(ptfem_mesh* m, int64_t* n_out, double* out_host, int64_t cap);
int coarse_finish_sums(ptfem_mesh* m, const double* sums_host, int64_t n);
// y_c = B_l Z_l^T r for every level; leaves r_c.y_c per level/system in cs.cdot
int coarse_apply(ptfem_ctx* ctx, CoarseSpace& cs, int S, const double* r);
// the same with the CG residual update fused into the restriction: r -= alpha q on every row as it is gathered, r.D^-1 r -> out_rz,
// r.r -> out_rr (device scalars [S]); partial: >= 8 * SMs * 2 * S doubles, ticket: a zeroed counter; after_restrict (optional) is
// recorded between the restriction and the grid hierarchy (fork point of work that may run beside the small grid kernels)
int coarse_apply_fused_update(ptfem_ctx* ctx, CoarseSpace& cs, int S, double* r, const double* q, const double* dinv, const double* alpha,
                              double* partial, unsigned int* ticket, double* out_rz, double* out_rr, cudaEvent_t after_restrict = nullptr);
CoarseDev coarse_dev(const CoarseSpace& cs);
void coarse_free(CoarseSpace* cs);
// row-partitioned solve (dist.cu): coarse spaces of a row block [row0, row0 + sys->nn) taken from the rank's replica
// of the mesh, restriction of the owned rows, and the replicated grid hierarchy (one right-hand side)
int coarse_replica_prepare(ptfem_mesh* full);   // api.cu
int coarse_attach_rows(ptfem_mesh* sys, ptfem_mesh* full, int64_t row0);
int coarse_restrict_rows(ptfem_ctx* ctx, CoarseSpace& cs, const double* r, double* rc_out);
int coarse_grids_apply(ptfem_ctx* ctx, CoarseSpace& cs, bool scaled0, int start = 0);
// sharded form of the two above (peer-memory transport): see coarse.cu
int coarse_touched_ranges(ptfem_ctx* ctx, CoarseSpace& cs, int64_t nn, int64_t ranges[4]);
int coarse_restrict_rows_sharded(ptfem_ctx* ctx, CoarseSpace& cs, const double* r, const int64_t ranges[4], double* out0, double* out1,
                                 const CoarseSignal& sig);
int coarse_prolong_finest_range(ptfem_ctx* ctx, CoarseSpace& cs, int64_t a0, int64_t b0);
}  // namespace ptfem
