/*
 * ptfem.h — C-ABI of libptfem.so, the B200-native steady-current-conduction FEM engine.
 *
 * Drop-in boundary.  The reference (alisabryantseva/pelvistim-fem) has no in-process
 * solver API: it shells out to `ElmerGrid 14 2 mesh.msh -out elmer_mesh` and
 * `ElmerSolver case.sif` (step03_ankle_layers/run_layered_sweep.py:1077,1099;
 * step04_pressure/run_pressure_sweep.py:696,727; step02_electrodes/run_sweep.py:315,327;
 * step01_box/test_step01_baseline.py:49,52) and reads results/case_t0001.vtu back.
 * Everything ElmerSolver does between "mesh directory + case.sif" and "nodal potential +
 * nodal volume current" is replaced by the entry points below; each one names the Elmer
 * stage / SIF keyword it stands in for.  Host code (Python, ctypes) parses the files and
 * calls these with plain pointers and sizes.
 *
 * Conventions: every function returns 0 on success, <0 on error (ptfem_last_error() gives the
 * message, thread-local).  Host buffers are caller-owned; device buffers are library-owned and
 * live behind the opaque handles.  Node / element indices are 0-based.  One context per GPU;
 * a handle is not thread-safe, different handles may be used from different threads.
 * All floating point is IEEE double.  There is no CPU fallback: without a CUDA device
 * ptfem_ctx_create fails.
 */
#ifndef PTFEM_H
#define PTFEM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ptfem_ctx ptfem_ctx;
typedef struct ptfem_mesh ptfem_mesh;

#define PTFEM_OK 0
#define PTFEM_ERR_ARG (-1)
#define PTFEM_ERR_CUDA (-2)
#define PTFEM_ERR_STATE (-3)
#define PTFEM_ERR_NOCONV (-4)
#define PTFEM_ERR_NCCL (-5)

/* preconditioners (replaces `Linear System Direct Method = UMFPACK`, step01_box/case.sif:41-42) */
#define PTFEM_PRECOND_JACOBI 0
#define PTFEM_PRECOND_CHEBYSHEV 1
#define PTFEM_PRECOND_TWOLEVEL 2 /* Jacobi + geometric coarse-grid correction: trilinear interpolation from regular grids over the
                                    bounding box, exact Galerkin inverse on the coarsest, diagonal (BPX) on finer ones; one shared
                                    matrix (multi-RHS) only */
#define PTFEM_PRECOND_AUTO (-1)  /* TWOLEVEL for one shared matrix with >= 100000 rows, else JACOBI */

/* nodal current recovery (replaces `Calculate Volume Current = True`, step01_box/case.sif:39) */
#define PTFEM_RECOVER_L2 0      /* Galerkin L2 projection, consistent mass matrix (PCG on the same pattern) */
#define PTFEM_RECOVER_LUMPED 1  /* volume-weighted nodal mean (row-sum lumped mass) */
#define PTFEM_RECOVER_AVERAGE 2 /* unweighted mean over the tets touching a node */

/* SpMV kernel variants for ptfem_spmv / ptfem_spmv_bench */
#define PTFEM_SPMV_AUTO 0
#define PTFEM_SPMV_VECTOR 1  /* sub-warp per row, direct global loads */
#define PTFEM_SPMV_STREAM 2  /* persistent CTAs, val/col staged into shared memory by bulk async copies, 2 stages */
#define PTFEM_SPMV_STREAM1 3 /* same, single stage (overlap across CTAs only) */

typedef struct ptfem_solve_opts {
  int32_t precond;      /* PTFEM_PRECOND_* */
  int32_t maxit;        /* iteration cap (per system) */
  int32_t check_every;  /* residual is copied to the host every this many iterations (0 = 50, or 10 with TWOLEVEL) */
  int32_t cheb_degree;  /* polynomial degree for PTFEM_PRECOND_CHEBYSHEV */
  double rtol;          /* stop when ||b - A x||_2 <= rtol * ||b||_2 for every system (default 1e-10), or when
                           residual replacement shows the round-off floor has been reached (still < 1e-8) */
  double cheb_ratio;    /* lambda_max / lambda_min assumed by the Chebyshev polynomial */
  int32_t spmv_variant; /* PTFEM_SPMV_* */
  int32_t use_graph;    /* capture check_every iterations in a CUDA graph */
  int32_t warm_start;   /* 0: start from phi = 0; 1: start from the solution already on the device */
  int32_t sample_spmv;  /* >0: after the solve, time this many launches of the solve's SpMV kernel (stats.spmv_ms) */
  int32_t coarse_nodes; /* PTFEM_PRECOND_TWOLEVEL: unknowns of the coarsest (exactly inverted) grid, 0 = default (300) */
  int32_t coarse_levels;/* PTFEM_PRECOND_TWOLEVEL: extra finer diagonal-only grids (0..5), each halves the cell size;
                           -1 (default) = as many as keep >= 16 mesh nodes per cell of the finest grid */
} ptfem_solve_opts;

typedef struct ptfem_solve_stats {
  int32_t iterations;   /* CG iterations executed (same for every system of the batch) */
  int32_t converged;    /* 1 if every system met rtol; 2 if the solve was accepted at the attainable accuracy instead (true residual
                           stopped falling under residual replacement, still <= 1e-8 but above rtol: see true_rel_residual); 0 otherwise */
  int32_t nsys;         /* systems solved at once */
  int32_t spmv_calls;   /* sparse matrix-vector products launched (all systems count as one) */
  double rel_residual;  /* max over systems of recurrence ||r|| / ||b|| at exit */
  double true_rel_residual; /* max over systems of ||b - A x|| / ||b|| recomputed at exit */
  double solve_ms;      /* device time of the iteration loop (CUDA events) */
  double spmv_ms;       /* device time of one SpMV launch, sampled with events after the solve */
  double setup_ms;      /* device time of the preconditioner set-up this solve paid (coarse grids, Galerkin inverse); 0 if reused */
  int32_t precond;      /* preconditioner actually used (PTFEM_PRECOND_AUTO resolved) */
  int32_t coarse_unknowns; /* unknowns of the exactly inverted coarse grid (0 without coarse space) */
} ptfem_solve_stats;

const char* ptfem_last_error(void);
int ptfem_version(void);
int ptfem_device_count(int* n);
/* PCI bus id of a device ("0000:1b:00.0", NUL-terminated, len >= 16): lets the host bind its threads and pinned staging
 * buffers to the NUMA node the GPU hangs off before it allocates them (engine.bind_host_to_gpu) */
int ptfem_device_pci_bus_id(int device, char* out, int len);

/* -- context: one per GPU ------------------------------------------------------------------ */
int ptfem_ctx_create(int device, ptfem_ctx** out);
int ptfem_ctx_destroy(ptfem_ctx* ctx);
int ptfem_ctx_sync(ptfem_ctx* ctx);
/* kernels launched by this context since creation (for bench.py's gpu_launches) */
int ptfem_ctx_launch_count(ptfem_ctx* ctx, int64_t* n);
/* the cudaStream_t every kernel of this context is launched on (for external CUDA-event timing) */
int ptfem_ctx_stream(ptfem_ctx* ctx, void** stream);

/* -- mesh: replaces ElmerSolver's reading of elmer_mesh/{mesh.nodes,mesh.elements,mesh.boundary}
 *    (formats: step01_box/find_boundaries.py:16-40,87-90).  Copies host arrays to the device. */
int ptfem_mesh_create(ptfem_ctx* ctx, int64_t nn, const double* xyz /*[nn*3]*/, int64_t nt,
                      const int32_t* tets /*[nt*4]*/, const int32_t* region /*[nt]*/, int64_t nb,
                      const int32_t* tris /*[nb*3]*/, const int32_t* bcid /*[nb]*/, ptfem_mesh** out);
/* same, but returns as soon as the copies are queued on an upload stream: the host arrays (pinned memory, or the call
 * degenerates to a blocking copy) must stay untouched until ptfem_pattern(), which waits for the upload and validates the
 * indices.  Lets a sweep upload the mesh of point k+1 while point k is being solved. */
int ptfem_mesh_create_async(ptfem_ctx* ctx, int64_t nn, const double* xyz /*[nn*3]*/, int64_t nt,
                            const int32_t* tets /*[nt*4]*/, const int32_t* region /*[nt]*/, int64_t nb,
                            const int32_t* tris /*[nb*3]*/, const int32_t* bcid /*[nb]*/, ptfem_mesh** out);
int ptfem_mesh_destroy(ptfem_mesh* m);
/* geometry change on fixed topology (node displacement; reference example:
 * run_layered_sweep.py:329-340): keeps the pattern, recomputes element geometry factors. */
int ptfem_mesh_set_coords(ptfem_mesh* m, const double* xyz /*[nn*3]*/);

/* -- K1: CSR pattern + element->nnz map (Elmer's matrix-structure creation). Idempotent. ----- */
int ptfem_pattern(ptfem_mesh* m, int64_t* nnz);
/* plan of the window SpMM (multi-RHS product out of shared-memory x windows) built with the pattern: info[0] = 1 if the numbering
 * allowed one (else the streaming kernel serves every product), [1] tiles, [2] largest window (vector rows), [3] largest tile blob
 * (bytes), [4] / [5] detected line length / plane size of the numbering; *rows_per_row = window rows staged per matrix row */
int ptfem_window_plan_info(ptfem_mesh* m, int64_t info[6], double* rows_per_row);
int ptfem_pattern_get(ptfem_mesh* m, int32_t* rowptr /*[nn+1]*/, int32_t* col /*[nnz]*/);
int ptfem_e2nnz_get(ptfem_mesh* m, int32_t* e2nnz /*[nt*16]*/);

/* -- K2+K3: bulk assembly of StatCurrentSolve (sum_e sigma_e V_e gradNi.gradNj).
 *    nsys > 1 assembles nsys matrices on the one pattern (batched-values mode): sigma is
 *    [nsys][nreg].  nsys == 1 and a later solve with nrhs > 1 is the multi-RHS mode. -------- */
int ptfem_assemble(ptfem_mesh* m, int32_t nreg, const int32_t* reg_ids /*[nreg]*/,
                   const double* sigma /*[nsys*nreg]*/, int32_t nsys);
int ptfem_values_get(ptfem_mesh* m, int32_t sys, int32_t with_bc, double* val /*[nnz]*/);

/* -- K4: boundary conditions (SIF `Boundary Condition` blocks).  rhs = -1 means every column. -- */
int ptfem_bc_reset(ptfem_mesh* m, int32_t nrhs);
/* `Potential = value` on every node of boundary elements with id bcid (step01_box/case.sif:61-71) */
int ptfem_bc_dirichlet(ptfem_mesh* m, int32_t rhs, int32_t bcid, double value);
/* `Current Density = g` on boundary elements with id bcid (run_layered_sweep.py:608-611);
 * positive g drives current into the domain */
int ptfem_bc_neumann(ptfem_mesh* m, int32_t rhs, int32_t bcid, double g);
/* same, for an explicit list of boundary-triangle indices (electrode patches of a sweep) */
int ptfem_bc_neumann_tris(ptfem_mesh* m, int32_t rhs, int64_t n, const int32_t* tri_idx, double g);
int ptfem_rhs_get(ptfem_mesh* m, int32_t rhs, double* b /*[nn]*/);

/* -- K5-K9: PCG solve (replaces the UMFPACK solve).  Systems solved at once =
 *    max(nsys of ptfem_assemble, nrhs of ptfem_bc_reset).  phi is [nsys_total][nn]. ---------- */
void ptfem_solve_opts_default(ptfem_solve_opts* o);
int ptfem_solve(ptfem_mesh* m, const ptfem_solve_opts* opts, double* phi, ptfem_solve_stats* stats);
/* device-resident variant: same work, no host copies (phi stays on the device for post-processing) */
int ptfem_solve_device(ptfem_mesh* m, const ptfem_solve_opts* opts, ptfem_solve_stats* stats);
int ptfem_phi_get(ptfem_mesh* m, int32_t sys, double* phi /*[nn]*/);
int ptfem_phi_set(ptfem_mesh* m, int32_t sys, const double* phi /*[nn]*/);
/* every system's potential, [nsys][nn], copied on the side stream and NOT waited for: phi (pinned host memory) is valid after
 * the next ptfem_ctx_sync (or ptfem_mesh_destroy); the copy overlaps the post-processing the caller enqueues next */
int ptfem_phi_get_all_async(ptfem_mesh* m, double* phi /*[nsys*nn], pinned*/);

/* y = A x with the assembled matrix of system sys (with_bc selects the eliminated matrix) */
int ptfem_spmv(ptfem_mesh* m, int32_t sys, int32_t with_bc, int32_t variant, const double* x, double* y);
/* time `iters` back-to-back SpMV launches of the given variant with CUDA events; nvec > 1 times
 * the multi-RHS / batched kernel on the current system layout */
int ptfem_spmv_bench(ptfem_mesh* m, int32_t variant, int32_t iters, double* ms_per_launch);

/* -- K10/K11: E = -grad phi, J = sigma E per element; nodal `volume current` ---------------- */
int ptfem_element_fields(ptfem_mesh* m, int32_t sys, double* E /*[nt*3] or NULL*/, double* J /*[nt*3] or NULL*/);
int ptfem_recover_current(ptfem_mesh* m, int32_t sys, int32_t method, double* J /*[nn*3] or NULL*/);

/* -- K12: metric reductions on the device fields of system sys (reference layer L5) ---------- */
typedef struct ptfem_footprint {
  double cx, cy, r;  /* centre and radius / half-side */
  int32_t square;    /* 0 disk, 1 square */
  int32_t pad_;
} ptfem_footprint;

/* over nodes with z > zmin, inside (mode 1) / outside both (mode 2) / ignoring (mode 0) the
 * footprints: out = {count, sum, max, min} of field: 0 |J|, 1 phi, 2 |J_z|, 3 J_z  */
int ptfem_metric_nodes(ptfem_mesh* m, int32_t sys, int32_t field, double zmin, double zmax, int32_t mode,
                       const ptfem_footprint* fp, int32_t nfp, double scale_r, double out[4]);
/* run_layered_sweep.py:704-761: sum J_z*area over boundary triangles whose centroid has
 * z > zmin and lies within scale_r * r of the footprint; out = {I_signed, area, count} */
int ptfem_metric_pad_current(ptfem_mesh* m, int32_t sys, double zmin, const ptfem_footprint* fp, double scale_r,
                             double out[3]);
/* run_layered_sweep.py:765-822,948-959: ROI sphere means over cells (tets then boundary tris) with the
 * VTK cell-data semantics; for each radius r0*mult[i]: out[i] = {n, sum|J_c|, sum|E_c|, n_z>z1, n_z in (z0,z1], n_z<=z0} */
int ptfem_metric_roi(ptfem_mesh* m, int32_t sys, const double cen[3], double r0, const double* mult, int32_t nmult,
                     double z0, double z1, int32_t include_tris, double* out /*[nmult*6]*/);
/* step01 (test_step01_baseline.py:59-104): centre-column least squares of phi(z):
 * out = {n, sum z, sum phi, sum z^2, sum z phi, sum phi^2} over nodes with hypot(x-cx,y-cy) < rad */
int ptfem_metric_column_fit(ptfem_mesh* m, int32_t sys, double cx, double cy, double rad, double out[6]);
/* moments of |J| over all nodes about `shift`: out = {n, sum(|J|-shift), sum((|J|-shift)^2)}
 * (two calls, shift = 0 then shift = mean, give mean and an accurate sample variance) */
int ptfem_metric_jstats(ptfem_mesh* m, int32_t sys, double shift, double out[3]);
/* weak-form reaction current: sum over nodes on boundary bcid of (K_raw phi - b_neumann)_i (exact KCL) */
int ptfem_metric_reaction(ptfem_mesh* m, int32_t sys, int32_t bcid, double* current);
/* K13 (not in the reference; north_star): phi sampled at points along a nerve-fibre polyline and the
 * activating function (second difference / h^2) at the interior points. */
int ptfem_sample_polyline(ptfem_mesh* m, int32_t sys, int64_t npts, const double* pts /*[npts*3]*/,
                          double* phi_out /*[npts]*/, double* af_out /*[npts]*/);
/* same, but the device->host copy of J runs on a side stream and is NOT waited for: J (pinned host memory) is valid
   after the next ptfem_ctx_sync; the following recovery waits for the copy before it overwrites the device buffer, so
   the copy overlaps the metric reductions of this system and the element pass of the next one */
int ptfem_recover_current_async(ptfem_mesh* m, int32_t sys, int32_t method, double* J /*[nn*3], pinned*/);
int ptfem_current_get(ptfem_mesh* m, double* J /*[nn*3]*/);
/* nodal currents of EVERY system of the last solve in two launches (lumped / average; L2 goes system by system):
 * the sweep form of `Calculate Volume Current` (one ElmerSolver run per sweep point in the reference,
 * run_sweep.py:301-341).  J, if not NULL, receives [nsys][nn][3]; wait = 0 (J in pinned memory) leaves the copy running
 * on the side stream until the next ptfem_ctx_sync.  Afterwards the per-system metric calls and
 * ptfem_recover_current(sys, same method) use these currents without recomputing them. */
int ptfem_recover_current_batch(ptfem_mesh* m, int32_t method, double* J /*[nsys*nn*3] or NULL*/, int32_t wait);

/* -- K12, batched: any number of the three reductions above, for any systems, in ONE pass over the mesh per kind and one
 *    device->host read-back (a sweep asks for the same few metrics of every configuration). ---------------------------- */
#define PTFEM_METRIC_NODES 0       /* ptfem_metric_nodes       -> out[0..3]  = {count, sum, max, min}             */
#define PTFEM_METRIC_PAD_CURRENT 1 /* ptfem_metric_pad_current -> out[0..2]  = {I_signed, area, count}            */
#define PTFEM_METRIC_ROI 2         /* ptfem_metric_roi         -> out[0..6*nmult-1]                               */
#define PTFEM_METRIC_OUT_STRIDE 24
typedef struct ptfem_metric_req {
  int32_t kind, sys;
  int32_t field, mode;      /* NODES */
  double zmin, zmax;        /* NODES, PAD_CURRENT (zmin) */
  double scale_r;           /* NODES, PAD_CURRENT */
  ptfem_footprint fp[2];    /* NODES (nfp of them), PAD_CURRENT (fp[0]) */
  int32_t nfp;
  int32_t include_tris;     /* ROI */
  double cen[3], r0, mult[4];
  int32_t nmult, pad_;
  double z0, z1;
} ptfem_metric_req;
int ptfem_metrics_batch(ptfem_mesh* m, int32_t nreq, const ptfem_metric_req* req, double* out /*[nreq][PTFEM_METRIC_OUT_STRIDE]*/);

/* -- multi-GPU: row-partitioned single solve (config #5).  The caller passes an ncclUniqueId
 *    (128 bytes) obtained on rank 0 via ptfem_dist_unique_id and broadcast by the launcher. ---- */
int ptfem_dist_unique_id(const char* libnccl_path, void* id128);
int ptfem_dist_init(ptfem_ctx* ctx, const char* libnccl_path, const void* id128, int32_t rank, int32_t nranks);
int ptfem_dist_finalize(ptfem_ctx* ctx);
/* local block of a row-partitioned system: rows [row0,row0+nloc) of the global matrix with columns
 * renumbered [0,nloc) = owned, [nloc,nloc+nhalo) = halo; for each neighbour rank the list of owned
 * rows to send and the halo slots to receive into. */
int ptfem_dist_system_create(ptfem_ctx* ctx, int64_t nloc, int64_t nhalo, const int32_t* rowptr, const int32_t* col,
                             const double* val, const double* b, int32_t nnbr, const int32_t* nbr_rank,
                             const int32_t* send_ptr, const int32_t* send_idx, const int32_t* recv_ptr,
                             ptfem_mesh** out);
int ptfem_dist_solve(ptfem_mesh* sys, const ptfem_solve_opts* opts, double* x_local, ptfem_solve_stats* stats,
                     double* ms_spmv, double* ms_halo, double* ms_allreduce);
/* Coarse-grid preconditioner for the row-partitioned solve.  `replica` is this rank's full mesh on the same context with
 * the same matrix assembled and boundary conditions set (the ranks build it anyway to cut their row blocks); its coarse
 * spaces are prepared here if they are not yet and the block [row0, row0+nloc) of them is attached to `sys`: grids and
 * Galerkin operators replicated, restriction / prolongation on the owned rows, the finest grid vector summed over the
 * ranks once per iteration.  The replica may be destroyed afterwards.  Call before ptfem_dist_p2p_export.
 * ptfem_dist_solve then preconditions with Jacobi + coarse grids for PTFEM_PRECOND_AUTO / _TWOLEVEL.
 * PTFEM_ERR_STATE: the Galerkin matrix is singular on this mesh (keep Jacobi). */
int ptfem_dist_coarse_attach(ptfem_mesh* sys, ptfem_mesh* replica, int64_t row0);
/* Distributed set-up (no rank holds the whole mesh): every rank creates a mesh of its OWNED nodes (global rows
 * [row0, row0+nloc), first), the ghost nodes of the elements touching them and the nodes of boundary triangles touching
 * those (partition.local_submesh), assembles and applies the boundary conditions on it - the owned rows of that matrix are
 * complete, with columns already numbered [owned | halo] - and builds its block with ptfem_dist_system_create from them.
 * ptfem_mesh_set_bbox gives the local mesh the bounding box of the WHOLE mesh and marks it as a part (a part may hold no
 * Dirichlet node: call it before the first solve / value read-back).  Coarse grids: ptfem_dist_coarse_partial
 * computes the Galerkin sums of the owned rows on grids chosen for nn_global nodes (sums == NULL: size query, *n doubles);
 * the launcher adds the arrays of all ranks; ptfem_dist_coarse_finish inverts them (identical on every rank); then
 * ptfem_dist_coarse_attach(sys, local_mesh, 0). */
int ptfem_mesh_set_bbox(ptfem_mesh* m, const double* lo /*[3]*/, const double* hi /*[3]*/);
int ptfem_dist_coarse_partial(ptfem_mesh* local_mesh, int64_t nrows_owned, int64_t nn_global, int32_t coarse_nodes,
                              int32_t coarse_levels, int64_t* n, double* sums /*[cap] or NULL*/, int64_t cap);
int ptfem_dist_coarse_finish(ptfem_mesh* local_mesh, const double* sums /*[n]*/, int64_t n);
/* Sharded coarse exchange of the peer-memory transport: a contiguous block of rows reaches a slab of the finest grid only.
 * _get returns this rank's {a0, b0, a1, b1} (finest-grid nodes [a0, b0) and level-1 nodes [a1, b1) its rows contribute to,
 * known after ptfem_dist_coarse_attach); the launcher all-gathers them and hands every rank the whole table with _set, before
 * ptfem_dist_p2p_connect.  Per iteration a rank then sums only the grid planes it shares with neighbouring slabs plus the
 * level-1 vector, instead of every rank's whole finest-grid vector.  Without _set the exchange covers the whole grids. */
int ptfem_dist_coarse_ranges_get(ptfem_mesh* sys, int64_t ranges4[4]);
int ptfem_dist_coarse_ranges_set(ptfem_mesh* sys, int32_t nranks, const int64_t* all_ranges /*[nranks*4]*/);
/* Peer-memory transport (NVLink P2P through CUDA IPC) instead of NCCL calls inside the iteration: every rank
 * exports two IPC handles (its vector and its mailbox, 2 x 64 bytes), the launcher all-gathers them, and each
 * rank connects.  halo_src[h] = index, in the owner's local numbering, of the row halo slot h mirrors.
 * After a successful connect ptfem_dist_solve pulls halos with direct peer loads and reduces the three CG
 * scalars through the mailboxes (ptfem_dist_init may then be called with id128 == NULL: no NCCL at all). */
int ptfem_dist_p2p_export(ptfem_mesh* sys, void* handles128);
int ptfem_dist_p2p_connect(ptfem_mesh* sys, int32_t nranks, const void* all_handles /*[nranks*128]*/,
                           const int32_t* halo_src /*[nhalo]*/);

#ifdef __cplusplus
}
#endif
#endif /* PTFEM_H */
