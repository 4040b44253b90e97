// api.cu — the extern "C" surface of libptfem.so (include/ptfem.h): argument checking, host<->device
// copies, and sequencing of the kernels in pattern.cu / assembly.cu / solver.cu / post.cu.
#include <algorithm>

#include "solver.cuh"

using namespace ptfem;

// internal entry points (other translation units)
int ptfem_build_pattern(ptfem_mesh* m);
int ptfem_build_geometry(ptfem_mesh* m);
int ptfem_do_assemble(ptfem_mesh* m, int32_t nreg, const int32_t* reg_ids, const double* sigma, int32_t nsys);
int ptfem_do_bc_reset(ptfem_mesh* m, int32_t nrhs);
int ptfem_do_bc_dirichlet(ptfem_mesh* m, int32_t rhs, int32_t bcid, double value);
int ptfem_do_bc_neumann(ptfem_mesh* m, int32_t rhs, int32_t bcid, double g);
int ptfem_do_bc_neumann_tris(ptfem_mesh* m, int32_t rhs, int64_t n, const int32_t* tri_idx, double g);
int ptfem_apply_bc(ptfem_mesh* m, double* dinv_out, int* S_out);
int ptfem_do_element_fields(ptfem_mesh* m, int sys);
int ptfem_do_recover(ptfem_mesh* m, int sys, int method);
int ptfem_do_recover_batch(ptfem_mesh* m, int method);
int ptfem_do_metrics_batch(ptfem_mesh* m, int32_t nreq, const ptfem_metric_req* req, double* out);
const double* ptfem_current_ptr(ptfem_mesh* m);
int ptfem_do_metric_nodes(ptfem_mesh* m, int sys, int field, double zmin, double zmax, int mode, const ptfem_footprint* fp,
                          int nfp, double scale_r, double out[4]);
int ptfem_do_metric_pad_current(ptfem_mesh* m, int sys, double zmin, const ptfem_footprint* fp, double scale_r, double out[3]);
int ptfem_do_metric_roi(ptfem_mesh* m, int sys, const double cen[3], double r0, const double* mult, int nmult, double z0,
                        double z1, int include_tris, double* out);
int ptfem_do_metric_column_fit(ptfem_mesh* m, int sys, double cx, double cy, double rad, double out[6]);
int ptfem_do_metric_jstats(ptfem_mesh* m, int sys, double shift, double out[3]);
int ptfem_do_metric_reaction(ptfem_mesh* m, int sys, int32_t bcid, double* current);
int ptfem_do_sample_polyline(ptfem_mesh* m, int sys, int64_t npts, const double* pts, double* phi_out, double* af_out);
void ptfem_dist_ctx_release(ptfem_ctx* ctx);
void ptfem_dist_mesh_release(ptfem_mesh* m);

namespace {

// strided gather/scatter between a [nn][S] interleaved device array and a dense [nn] host-bound staging vector
__global__ void gather_sys_kernel(const double* __restrict__ in, int64_t nn, int S, int sys, double* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nn) out[i] = in[i * S + sys];
}
__global__ void scatter_sys_kernel(const double* __restrict__ in, int64_t nn, int S, int sys, double* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nn) out[i * S + sys] = in[i];
}
// [nn][S] -> [nsys][nn] (system-major, what the host API hands back)
__global__ void transpose_out_kernel(const double* __restrict__ in, int64_t nn, int S, int nsys, double* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  for (int s = 0; s < nsys; ++s) out[(int64_t)s * nn + i] = in[i * S + s];
}

// first out-of-range entry of an index array (position, or INT64_MAX) — mesh validation on the device
__global__ void index_check_kernel(const int32_t* __restrict__ idx, int64_t n, int64_t nn, unsigned long long* __restrict__ first_bad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && (idx[i] < 0 || idx[i] >= nn)) atomicMin(first_bad, (unsigned long long)i);
}
// per-CTA bounding box partials [grid][6] (lo xyz, hi xyz); the host folds the few hundred partials
__global__ void __launch_bounds__(256) bbox_kernel(const double* __restrict__ xyz, int64_t nn, double* __restrict__ part) {
  __shared__ double s[6][256];
  double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < nn; i += (int64_t)gridDim.x * 256)
    for (int d = 0; d < 3; ++d) {
      const double v = xyz[3 * i + d];
      lo[d] = fmin(lo[d], v);
      hi[d] = fmax(hi[d], v);
    }
  for (int d = 0; d < 3; ++d) {
    s[d][threadIdx.x] = lo[d];
    s[3 + d][threadIdx.x] = hi[d];
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double v = s[threadIdx.x][0];
    for (int k = 1; k < 256; ++k) v = threadIdx.x < 3 ? fmin(v, s[threadIdx.x][k]) : fmax(v, s[threadIdx.x][k]);
    part[(size_t)blockIdx.x * 6 + threadIdx.x] = v;
  }
}

int device_bbox(ptfem_mesh* m) {
  ptfem_ctx* ctx = m->ctx;
  const int grid = std::min<int64_t>(ceil_div(m->nn, 256), 4 * ctx->sm_count);
  DevBuf<double> part;
  PT_TRY(part.alloc((size_t)grid * 6));
  bbox_kernel<<<grid, 256, 0, ctx->stream>>>(m->xyz.p, m->nn, part.p);
  PT_LAUNCH_CHECK(ctx);
  std::vector<double> h((size_t)grid * 6);
  PT_CK(cudaMemcpyAsync(h.data(), part.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  PT_CK(cudaStreamSynchronize(ctx->stream));
  for (int d = 0; d < 3; ++d) {
    m->bb_lo[d] = h[d];
    m->bb_hi[d] = h[3 + d];
  }
  for (int b = 1; b < grid; ++b)
    for (int d = 0; d < 3; ++d) {
      m->bb_lo[d] = std::min(m->bb_lo[d], h[(size_t)b * 6 + d]);
      m->bb_hi[d] = std::max(m->bb_hi[d], h[(size_t)b * 6 + 3 + d]);
    }
  return PTFEM_OK;
}

int make_linsys(ptfem_mesh* m, LinSys& A) {
  A.nn = m->nn;
  A.nnz = m->nnz;
  A.rowptr = m->rowptr.p;
  A.col = m->col.p;
  A.val = m->val_bc.p;
  A.VS = m->nvalp;
  A.S = m->S;
  A.dinv = m->dinv.p;
  A.b = m->b.p;
  A.stream_rows = m->stream_rows;
  A.stream_cap = m->stream_cap;
  if (m->has_qcopy && m->nvalp == 1 && m->qval.p) {
    A.qrowptr = m->qrowptr.p;
    A.qcol = m->qcol.p;
    A.qval = m->qval.p;
    A.q_rows = m->q_rows;
    A.q_cap = m->q_cap;
  }
  if (m->win.valid && m->nvalp == 1) A.win = &m->win;
  if (m->has_rowperm && m->nvalp == 1 && m->pval.p) {
    A.rowid = m->rowid.p;
    A.prowptr = m->prowptr.p;
    A.pcol = m->pcol.p;
    A.pval = m->pval.p;
  }
  return PTFEM_OK;
}

int prepare_systems(ptfem_mesh* m) {
  if (!m->has_pattern) return set_err(PTFEM_ERR_STATE, "ptfem_pattern has not been called");
  if (m->bc_dirty || !m->b.p) {
    PT_TRY(m->dinv.alloc((size_t)m->nn * (m->nvalp > 0 ? m->nvalp : 1)));
    int S = 0;
    PT_TRY(ptfem_apply_bc(m, m->dinv.p, &S));
    if (S != m->S || !m->phi.p) {
      m->S = S;
      PT_TRY(m->phi.alloc((size_t)m->nn * S));
      PT_CK(cudaMemsetAsync(m->phi.p, 0, (size_t)m->nn * S * sizeof(double), m->ctx->stream));
    }
    m->nsys_user = std::max(m->nval, m->nrhs);
    if (m->has_qcopy && m->nvalp == 1) {     // even-padded copy of the eliminated matrix for the multi-RHS streaming kernel
      PT_TRY(m->qval.alloc(m->qnnz + 8));
      PT_CK(cudaMemsetAsync(m->qval.p + m->qnnz, 0, 8 * sizeof(double), m->ctx->stream));
      PT_TRY(pad_values(m->ctx, m->nn, m->rowptr.p, m->qrowptr.p, m->val_bc.p, m->qval.p));
    }
    if (m->win.valid && m->nvalp == 1) PT_TRY(window_refresh_values(m, m->val_bc.p));   // the window SpMM's blobs
    if (m->has_rowperm && m->nvalp == 1) {   // the streaming kernel's private copy of the eliminated matrix
      PT_TRY(m->pval.alloc(m->nnz + 8));
      PT_CK(cudaMemsetAsync(m->pval.p + m->nnz, 0, 8 * sizeof(double), m->ctx->stream));
      PT_TRY(permute_values(m->ctx, m->nn, m->rowptr.p, m->rowid.p, m->prowptr.p, m->val_bc.p, m->pval.p));
    }
  }
  return PTFEM_OK;
}

}  // namespace

namespace ptfem {
// coarse spaces of a rank's replica of the mesh for the row-partitioned solve (dist.cu): the defaults of
// ptfem_solve's automatic choice, including its retry on a coarser grid
int coarse_replica_prepare(ptfem_mesh* full) {
  PT_TRY(prepare_systems(full));
  if (full->nvalp != 1) return set_err(PTFEM_ERR_ARG, "the replica must hold one matrix");
  // spaces prepared by an earlier solve of the replica on this matrix are taken as they are (the caller chose them)
  if (full->coarse && full->coarse->geom_ok && full->coarse->matrix_epoch == full->matrix_epoch) return PTFEM_OK;
  int rc = coarse_prepare(full, 0, -1, full->S);
  if (rc == PTFEM_ERR_STATE) rc = coarse_prepare(full, 80, -1, full->S);
  return rc;
}
}  // namespace ptfem

extern "C" {

const char* ptfem_last_error(void) { return g_err.c_str(); }
int ptfem_version(void) { return 100; }

int ptfem_device_count(int* n) {
  PT_ARG(n, "null pointer");
  cudaError_t e = cudaGetDeviceCount(n);
  if (e != cudaSuccess) {
    *n = 0;
    return set_err(PTFEM_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
  }
  return PTFEM_OK;
}

int ptfem_device_pci_bus_id(int device, char* out, int len) {
  PT_ARG(out && len >= 16, "buffer of at least 16 bytes needed");
  out[0] = 0;
  PT_CK(cudaDeviceGetPCIBusId(out, len, device));
  return PTFEM_OK;
}

int ptfem_ctx_create(int device, ptfem_ctx** out) {
  PT_ARG(out, "null pointer");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return set_err(PTFEM_ERR_CUDA, "no CUDA device available (%s); libptfem has no CPU fallback",
                   e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  if (device < 0 || device >= n) return set_err(PTFEM_ERR_ARG, "device %d out of range (0..%d)", device, n - 1);
  PT_CK(cudaSetDevice(device));
  ptfem_ctx* c = new ptfem_ctx();
  c->device = device;
  cudaDeviceProp prop;
  PT_CK(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  PT_CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  PT_CK(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
  PT_CK(cudaStreamCreateWithFlags(&c->stream3, cudaStreamNonBlocking));
  PT_CK(cudaStreamCreateWithFlags(&c->stream_x, cudaStreamNonBlocking));
  PT_CK(cudaEventCreateWithFlags(&c->ev_xfork, cudaEventDisableTiming));
  PT_CK(cudaEventCreateWithFlags(&c->ev_xjoin, cudaEventDisableTiming));
  PT_CK(cudaEventCreateWithFlags(&c->ev_j_ready, cudaEventDisableTiming));
  PT_CK(cudaEventCreateWithFlags(&c->ev_j_copied, cudaEventDisableTiming));
  PT_CK(cudaEventCreateWithFlags(&c->ev_phi_ready, cudaEventDisableTiming));
  if (const char* e = getenv("PTFEM_INTERLEAVE")) c->tune_interleave = atoi(e) != 0;
  if (const char* e = getenv("PTFEM_MORTON")) c->tune_morton = atoi(e);
  if (const char* e = getenv("PTFEM_P2P_FUSED")) c->tune_p2p_fused = atoi(e) != 0;
  if (const char* e = getenv("PTFEM_XPREFETCH")) c->tune_xprefetch = atoi(e) != 0;
  if (const char* e = getenv("PTFEM_CTAS_PER_SM")) c->tune_ctas_per_sm = atoi(e);
  if (const char* e = getenv("PTFEM_RESTRICT_OCC")) c->tune_restrict_occ = atoi(e);
  if (const char* e = getenv("PTFEM_COARSE_FUSED")) c->tune_coarse_fused = atoi(e) != 0;
  if (const char* e = getenv("PTFEM_PUPDATE_NP")) c->tune_pupdate_np = atoi(e) == 2 ? 2 : 1;
  if (const char* e = getenv("PTFEM_PUPDATE_OCC")) c->tune_pupdate_occ = atoi(e);
  if (const char* e = getenv("PTFEM_PUPDATE_GRID")) c->tune_pupdate_grid = atoi(e);
  if (const char* e = getenv("PTFEM_CHAIN_TAIL")) c->tune_chain_tail = atoi(e) != 0;
  if (const char* e = getenv("PTFEM_SPMM_PAIR")) c->tune_pair = atoi(e) != 0;
  if (const char* e = getenv("PTFEM_COARSE_WEIGHT")) {
    const double w = atof(e);
    if (w > 0.0) c->tune_coarse_weight = w;
  }
  if (const char* e = getenv("PTFEM_FUSE_UPDATE")) c->tune_fuse_update = atoi(e);
  if (const char* e = getenv("PTFEM_FUSE_OCC")) c->tune_fuse_occ = atoi(e);
  if (const char* e = getenv("PTFEM_FUSE_GRID")) c->tune_fuse_grid = atoi(e);
  if (const char* e = getenv("PTFEM_FUSE_PREFETCH")) c->tune_fuse_prefetch = atoi(e);
  if (const char* e = getenv("PTFEM_FUSE_PIPE")) c->tune_fuse_pipe = atoi(e);
  if (const char* e = getenv("PTFEM_RESTRICT_GRID")) c->tune_restrict_grid = atoi(e);
  if (const char* e = getenv("PTFEM_SPLIT_X")) c->tune_split_x = atoi(e);
  if (const char* e = getenv("PTFEM_SPLIT_X_CTAS")) c->tune_split_x_ctas = atoi(e);
  if (const char* e = getenv("PTFEM_SPMM_WINDOW")) c->tune_window = atoi(e);
  if (const char* e = getenv("PTFEM_WINDOW_BX")) c->tune_window_bx = atoi(e);
  if (const char* e = getenv("PTFEM_WINDOW_CTAS")) c->tune_window_ctas = atoi(e);
  if (const char* e = getenv("PTFEM_STREAM_CAP")) c->tune_stream_cap = atoi(e);
  if (const char* e = getenv("PTFEM_STREAM_ROWS")) c->tune_stream_rows = atoi(e);
  if (const char* e = getenv("PTFEM_STREAM_TPR")) c->tune_stream_tpr = atoi(e);
  if (const char* e = getenv("PTFEM_STREAM_STAGES")) c->tune_stream_stages = atoi(e);
  c->h_pinned_n = 4096;
  PT_CK(cudaMallocHost((void**)&c->h_pinned, c->h_pinned_n * sizeof(double)));
  *out = c;
  return PTFEM_OK;
}

int ptfem_ctx_destroy(ptfem_ctx* ctx) {
  if (!ctx) return PTFEM_OK;
  cudaSetDevice(ctx->device);
  ptfem_dist_ctx_release(ctx);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->ev_j_ready) cudaEventDestroy(ctx->ev_j_ready);
  if (ctx->ev_j_copied) cudaEventDestroy(ctx->ev_j_copied);
  if (ctx->ev_phi_ready) cudaEventDestroy(ctx->ev_phi_ready);
  if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
  if (ctx->stream3) cudaStreamDestroy(ctx->stream3);
  if (ctx->stream_x) cudaStreamDestroy(ctx->stream_x);
  if (ctx->ev_xfork) cudaEventDestroy(ctx->ev_xfork);
  if (ctx->ev_xjoin) cudaEventDestroy(ctx->ev_xjoin);
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  delete ctx;
  dev_cache_flush();
  return PTFEM_OK;
}

int ptfem_ctx_sync(ptfem_ctx* ctx) {
  PT_ARG(ctx, "null context");
  PT_CK(cudaSetDevice(ctx->device));
  PT_CK(cudaStreamSynchronize(ctx->stream));
  PT_CK(cudaStreamSynchronize(ctx->stream2));
  return PTFEM_OK;
}

int ptfem_ctx_launch_count(ptfem_ctx* ctx, int64_t* n) {
  PT_ARG(ctx && n, "null pointer");
  *n = ctx->launches;
  return PTFEM_OK;
}

int ptfem_ctx_stream(ptfem_ctx* ctx, void** stream) {
  PT_ARG(ctx && stream, "null pointer");
  *stream = (void*)ctx->stream;
  return PTFEM_OK;
}

namespace {
// device-side validation of the uploaded index arrays + bounding box; ends an asynchronous upload
int mesh_finish_upload(ptfem_mesh* m) {
  if (!m->upload_pending) return PTFEM_OK;
  ptfem_ctx* ctx = m->ctx;
  m->upload_pending = false;
  PT_CK(cudaStreamWaitEvent(ctx->stream, m->ev_upload, 0));
  // (a host pass over 80 M indices cost more than the upload itself)
  DevBuf<unsigned long long> bad;
  PT_TRY(bad.alloc(2));
  PT_CK(cudaMemsetAsync(bad.p, 0xff, 2 * sizeof(unsigned long long), ctx->stream));
  if (m->nt > 0) index_check_kernel<<<ceil_div(m->nt * 4, 256), 256, 0, ctx->stream>>>(m->tets.p, m->nt * 4, m->nn, bad.p);
  if (m->nb > 0) index_check_kernel<<<ceil_div(m->nb * 3, 256), 256, 0, ctx->stream>>>(m->tris.p, m->nb * 3, m->nn, bad.p + 1);
  ctx->launches += (m->nt > 0) + (m->nb > 0);
  unsigned long long hb[2];
  PT_CK(cudaMemcpyAsync(hb, bad.p, sizeof hb, cudaMemcpyDeviceToHost, ctx->stream));
  PT_CK(cudaStreamSynchronize(ctx->stream));
  if (hb[0] != ~0ull) return set_err(PTFEM_ERR_ARG, "tet %llu refers to a node outside 0..%lld", hb[0] / 4, (long long)m->nn - 1);
  if (hb[1] != ~0ull) return set_err(PTFEM_ERR_ARG, "boundary triangle %llu refers to a node outside 0..%lld", hb[1] / 3, (long long)m->nn - 1);
  return device_bbox(m);
}
}  // namespace

int ptfem_mesh_create_async(ptfem_ctx* ctx, int64_t nn, const double* xyz, int64_t nt, const int32_t* tets, const int32_t* region,
                            int64_t nb, const int32_t* tris, const int32_t* bcid, ptfem_mesh** out) {
  PT_ARG(ctx && out, "null pointer");
  *out = nullptr;
  PT_ARG(nn > 0 && nt >= 0 && nb >= 0, "negative or zero sizes");
  PT_ARG(xyz && (nt == 0 || (tets && region)) && (nb == 0 || (tris && bcid)), "null array");
  PT_ARG(nn < 2147483647LL && nt * 16 < 2147483647LL * 4, "mesh too large for 32-bit indexing");
  PT_CK(cudaSetDevice(ctx->device));
  ptfem_mesh* m = new ptfem_mesh();
  m->ctx = ctx;
  m->nn = nn;
  m->nt = nt;
  m->nb = nb;
  int rc = PTFEM_OK;
  if (cudaEventCreateWithFlags(&m->ev_upload, cudaEventDisableTiming) != cudaSuccess) rc = set_err(PTFEM_ERR_CUDA, "event creation failed");
  auto up = [&](auto& buf, const auto* src, size_t count) {
    if (rc) return;
    rc = buf.alloc(count);
    if (rc || count == 0) return;
    cudaError_t e = cudaMemcpyAsync(buf.p, src, count * sizeof(*src), cudaMemcpyHostToDevice, ctx->stream3);
    if (e != cudaSuccess) rc = set_err(PTFEM_ERR_CUDA, "mesh upload: %s", cudaGetErrorString(e));
  };
  up(m->tets, tets, (size_t)nt * 4);       // what the pattern needs first
  up(m->xyz, xyz, (size_t)nn * 3);
  up(m->region, region, (size_t)nt);
  up(m->tris, tris, (size_t)nb * 3);
  up(m->bcid, bcid, (size_t)nb);
  if (!rc && cudaEventRecord(m->ev_upload, ctx->stream3) != cudaSuccess) rc = set_err(PTFEM_ERR_CUDA, "mesh upload failed");
  if (rc) {
    cudaStreamSynchronize(ctx->stream3);
    if (m->ev_upload) cudaEventDestroy(m->ev_upload);
    delete m;
    return rc;
  }
  m->upload_pending = true;
  *out = m;
  return PTFEM_OK;
}

int ptfem_mesh_create(ptfem_ctx* ctx, int64_t nn, const double* xyz, int64_t nt, const int32_t* tets, const int32_t* region,
                      int64_t nb, const int32_t* tris, const int32_t* bcid, ptfem_mesh** out) {
  PT_TRY(ptfem_mesh_create_async(ctx, nn, xyz, nt, tets, region, nb, tris, bcid, out));
  PT_CK(cudaStreamSynchronize(ctx->stream3));       // the host arrays are the caller's again
  const int rc = mesh_finish_upload(*out);
  if (rc) {
    ptfem_mesh_destroy(*out);
    *out = nullptr;
  }
  return rc;
}

int ptfem_mesh_destroy(ptfem_mesh* m) {
  if (!m) return PTFEM_OK;
  cudaSetDevice(m->ctx->device);
  cudaStreamSynchronize(m->ctx->stream);
  cudaStreamSynchronize(m->ctx->stream2);
  if (m->upload_pending) cudaStreamSynchronize(m->ctx->stream3);
  if (m->ev_upload) cudaEventDestroy(m->ev_upload);
  pcg_work_drop_graph(m->work);
  pcg_work_drop_graph(m->work3);
  ptfem_dist_mesh_release(m);
  if (m->coarse) coarse_free(m->coarse);
  m->coarse = nullptr;
  delete m;
  return PTFEM_OK;
}

int ptfem_mesh_set_coords(ptfem_mesh* m, const double* xyz) {
  PT_ARG(m && xyz, "null pointer");
  PT_CK(cudaSetDevice(m->ctx->device));
  PT_TRY(mesh_finish_upload(m));
  PT_CK(cudaMemcpyAsync(m->xyz.p, xyz, (size_t)m->nn * 3 * sizeof(double), cudaMemcpyHostToDevice, m->ctx->stream));
  PT_CK(cudaStreamSynchronize(m->ctx->stream));
  PT_TRY(device_bbox(m));
  if (m->coarse) m->coarse->geom_ok = false;
  m->has_geom = false;
  m->has_tcen = false;
  if (m->has_pattern) PT_TRY(ptfem_build_geometry(m));
  m->nval = 0;  // values must be re-assembled
  m->bc_dirty = true;
  return PTFEM_OK;
}

int ptfem_mesh_set_bbox(ptfem_mesh* m, const double* lo, const double* hi) {
  PT_ARG(m && lo && hi, "null pointer");
  for (int d = 0; d < 3; ++d) {
    PT_ARG(hi[d] >= lo[d], "empty bounding box");
    m->bb_lo[d] = lo[d];
    m->bb_hi[d] = hi[d];
  }
  if (m->coarse) m->coarse->geom_ok = false;
  m->is_part = true;
  return PTFEM_OK;
}

int ptfem_dist_coarse_partial(ptfem_mesh* m, int64_t nrows_owned, int64_t nn_global, int32_t coarse_nodes, int32_t coarse_levels,
                              int64_t* n, double* sums, int64_t cap) {
  PT_ARG(m && !m->is_dist && n, "needs the rank's local mesh");
  PT_ARG(nrows_owned > 0 && nrows_owned <= m->nn && nn_global >= nrows_owned, "bad row counts");
  PT_CK(cudaSetDevice(m->ctx->device));
  PT_TRY(prepare_systems(m));
  if (m->nvalp != 1 || m->S != 1) return set_err(PTFEM_ERR_ARG, "the distributed set-up handles one matrix, one right-hand side");
  if (!m->coarse || !m->coarse->partial_ready || m->coarse->row_limit != nrows_owned || m->coarse->nn_levels != nn_global) {
    if (m->coarse) coarse_free(m->coarse);
    m->coarse = new CoarseSpace();
    m->coarse->row_limit = nrows_owned;
    m->coarse->nn_levels = nn_global;
    m->coarse->partial_mode = true;
    PT_TRY(coarse_prepare(m, coarse_nodes, coarse_levels, 1));
  }
  return coarse_partial_sums(m, n, sums, cap);
}

int ptfem_dist_coarse_finish(ptfem_mesh* m, const double* sums, int64_t n) {
  PT_ARG(m && !m->is_dist && sums, "needs the rank's local mesh and the summed Galerkin data");
  PT_CK(cudaSetDevice(m->ctx->device));
  return coarse_finish_sums(m, sums, n);
}

int ptfem_window_plan_info(ptfem_mesh* m, int64_t info[6], double* rows_per_row) {
  PT_ARG(m && info && rows_per_row, "null pointer");
  if (!m->has_pattern) return set_err(PTFEM_ERR_STATE, "ptfem_pattern has not been called");
  info[0] = m->win.valid ? 1 : 0;
  info[1] = m->win.ntiles;
  info[2] = m->win.wmax;
  info[3] = m->win.capblob;
  info[4] = m->win.grid_a;
  info[5] = m->win.grid_b;
  *rows_per_row = m->win.window_rows_per_row;
  return PTFEM_OK;
}

int ptfem_pattern(ptfem_mesh* m, int64_t* nnz) {
  PT_ARG(m, "null mesh");
  PT_CK(cudaSetDevice(m->ctx->device));
  PT_TRY(mesh_finish_upload(m));
  if (!m->has_pattern) PT_TRY(ptfem_build_pattern(m));
  if (!m->has_geom) PT_TRY(ptfem_build_geometry(m));
  if (nnz) *nnz = m->nnz;
  return PTFEM_OK;
}

int ptfem_pattern_get(ptfem_mesh* m, int32_t* rowptr, int32_t* col) {
  PT_ARG(m, "null mesh");
  if (!m->has_pattern) return set_err(PTFEM_ERR_STATE, "ptfem_pattern has not been called");
  PT_CK(cudaSetDevice(m->ctx->device));
  if (rowptr) PT_CK(cudaMemcpyAsync(rowptr, m->rowptr.p, (size_t)(m->nn + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, m->ctx->stream));
  if (col) PT_CK(cudaMemcpyAsync(col, m->col.p, (size_t)m->nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, m->ctx->stream));
  PT_CK(cudaStreamSynchronize(m->ctx->stream));
  return PTFEM_OK;
}

int ptfem_e2nnz_get(ptfem_mesh* m, int32_t* e2nnz) {
  PT_ARG(m && e2nnz, "null pointer");
  if (!m->has_pattern) return set_err(PTFEM_ERR_STATE, "ptfem_pattern has not been called");
  PT_CK(cudaSetDevice(m->ctx->device));
  PT_CK(cudaMemcpyAsync(e2nnz, m->e2nnz.p, (size_t)m->nt * 16 * sizeof(int32_t), cudaMemcpyDeviceToHost, m->ctx->stream));
  PT_CK(cudaStreamSynchronize(m->ctx->stream));
  return PTFEM_OK;
}

int ptfem_assemble(ptfem_mesh* m, int32_t nreg, const int32_t* reg_ids, const double* sigma, int32_t nsys) {
  PT_ARG(m && reg_ids && sigma, "null pointer");
  PT_ARG(nsys >= 1, "nsys must be >= 1");
  PT_CK(cudaSetDevice(m->ctx->device));
  PT_TRY(ptfem_pattern(m, nullptr));
  return ptfem_do_assemble(m, nreg, reg_ids, sigma, nsys);
}

int ptfem_values_get(ptfem_mesh* m, int32_t sys, int32_t with_bc, double* val) {
  PT_ARG(m && val, "null pointer");
  if (m->nval < 1) return set_err(PTFEM_ERR_STATE, "ptfem_assemble has not been called");
  PT_ARG(sys >= 0 && sys < m->nval, "system index out of range");
  PT_CK(cudaSetDevice(m->ctx->device));
  if (with_bc) PT_TRY(prepare_systems(m));
  const double* src = with_bc ? m->val_bc.p : m->val_raw.p;
  DevBuf<double> tmp;
  PT_TRY(tmp.alloc(m->nnz));
  gather_sys_kernel<<<ceil_div(m->nnz, 256), 256, 0, m->ctx->stream>>>(src, m->nnz, m->nvalp, sys, tmp.p);
  PT_LAUNCH_CHECK(m->ctx);
  PT_CK(cudaMemcpyAsync(val, tmp.p, (size_t)m->nnz * sizeof(double), cudaMemcpyDeviceToHost, m->ctx->stream));
  PT_CK(cudaStreamSynchronize(m->ctx->stream));
  return PTFEM_OK;
}

int ptfem_bc_reset(ptfem_mesh* m, int32_t nrhs) {
  PT_ARG(m, "null mesh");
  PT_ARG(nrhs >= 1, "nrhs must be >= 1");
  PT_CK(cudaSetDevice(m->ctx->device));
  PT_TRY(ptfem_pattern(m, nullptr));
  return ptfem_do_bc_reset(m, nrhs);
}

#define PT_RHS_CHECK()                                                              \
  PT_ARG(m, "null mesh");                                                           \
  if (m->nrhs < 1) return set_err(PTFEM_ERR_STATE, "ptfem_bc_reset has not been called"); \
  PT_ARG(rhs >= -1 && rhs < m->nrhs, "right-hand-side index out of range");         \
  PT_CK(cudaSetDevice(m->ctx->device))

int ptfem_bc_dirichlet(ptfem_mesh* m, int32_t rhs, int32_t bcid, double value) {
  PT_RHS_CHECK();
  return ptfem_do_bc_dirichlet(m, rhs, bcid, value);
}
int ptfem_bc_neumann(ptfem_mesh* m, int32_t rhs, int32_t bcid, double g) {
  PT_RHS_CHECK();
  return ptfem_do_bc_neumann(m, rhs, bcid, g);
}
int ptfem_bc_neumann_tris(ptfem_mesh* m, int32_t rhs, int64_t n, const int32_t* tri_idx, double g) {
  PT_RHS_CHECK();
  PT_ARG(n == 0 || tri_idx, "null pointer");
  return ptfem_do_bc_neumann_tris(m, rhs, n, tri_idx, g);
}

int ptfem_rhs_get(ptfem_mesh* m, int32_t rhs, double* b) {
  PT_ARG(m && b, "null pointer");
  PT_CK(cudaSetDevice(m->ctx->device));
  PT_TRY(prepare_systems(m));
  PT_ARG(rhs >= 0 && rhs < m->nsys_user, "right-hand-side index out of range");
  DevBuf<double> tmp;
  PT_TRY(tmp.alloc(m->nn));
  gather_sys_kernel<<<ceil_div(m->nn, 256), 256, 0, m->ctx->stream>>>(m->b.p, m->nn, m->S, rhs, tmp.p);
  PT_LAUNCH_CHECK(m->ctx);
  PT_CK(cudaMemcpyAsync(b, tmp.p, (size_t)m->nn * sizeof(double), cudaMemcpyDeviceToHost, m->ctx->stream));
  PT_CK(cudaStreamSynchronize(m->ctx->stream));
  return PTFEM_OK;
}

void ptfem_solve_opts_default(ptfem_solve_opts* o) {
  if (!o) return;
  o->precond = PTFEM_PRECOND_AUTO;
  o->maxit = 200000;
  o->check_every = 0;
  o->cheb_degree = 4;
  o->rtol = 1e-10;
  o->cheb_ratio = 30.0;
  o->spmv_variant = PTFEM_SPMV_AUTO;
  o->use_graph = 1;
  o->warm_start = 0;
  o->sample_spmv = 0;
  o->coarse_nodes = 0;
  o->coarse_levels = -1;
}

int ptfem_solve_device(ptfem_mesh* m, const ptfem_solve_opts* opts, ptfem_solve_stats* stats) {
  PT_ARG(m, "null mesh");
  PT_CK(cudaSetDevice(m->ctx->device));
  ptfem_solve_opts o;
  if (opts) o = *opts; else ptfem_solve_opts_default(&o);
  PT_ARG(o.rtol > 0.0 && o.maxit > 0, "rtol and maxit must be positive");
  PT_ARG(o.precond >= PTFEM_PRECOND_AUTO && o.precond <= PTFEM_PRECOND_TWOLEVEL, "unknown preconditioner");
  PT_TRY(prepare_systems(m));
  const bool automatic = o.precond == PTFEM_PRECOND_AUTO;
  if (automatic) o.precond = m->nn >= 100000 ? PTFEM_PRECOND_TWOLEVEL : PTFEM_PRECOND_JACOBI;   // shared and batched matrices alike
  LinSys A;
  make_linsys(m, A);
  double setup_ms = 0.0;
  if (o.precond == PTFEM_PRECOND_TWOLEVEL) {
    const int64_t epoch_before = m->coarse ? m->coarse->matrix_epoch : -1;
    int rcp = coarse_prepare(m, o.coarse_nodes, o.coarse_levels, m->S);
    // a strongly graded mesh can leave coarse cells (nearly) empty and the Galerkin matrix singular: the automatic
    // choice then retries with a 4x coarser grid and finally settles for Jacobi; an explicit request reports the error
    if (rcp == PTFEM_ERR_STATE && automatic) {
      const int nodes = (o.coarse_nodes > 0 ? o.coarse_nodes : kDefaultCoarseNodes) / 4;
      rcp = coarse_prepare(m, nodes < 27 ? 27 : nodes, o.coarse_levels, m->S);
      if (rcp == PTFEM_ERR_STATE) {
        o.precond = PTFEM_PRECOND_JACOBI;
        rcp = PTFEM_OK;
      }
    }
    if (rcp) return rcp;
    if (o.precond == PTFEM_PRECOND_TWOLEVEL) {
      if (m->coarse->matrix_epoch != epoch_before) setup_ms = m->coarse->setup_ms;
      A.coarse = m->coarse;
    }
  }
  m->J_sys = -1;
  m->J_all_valid = false;
  m->J_from_all = false;
  if (!o.warm_start) PT_CK(cudaMemsetAsync(m->phi.p, 0, (size_t)m->nn * m->S * sizeof(double), m->ctx->stream));
  int rc = pcg_solve(m->ctx, A, m->work, o, m->phi.p, stats);
  if (stats) {
    stats->nsys = m->nsys_user;
    stats->setup_ms = setup_ms;
    stats->precond = o.precond;
    stats->coarse_unknowns = (o.precond == PTFEM_PRECOND_TWOLEVEL) ? (int32_t)m->coarse->lev[m->coarse->nlev - 1].k : 0;
  }
  return rc;
}

int ptfem_solve(ptfem_mesh* m, const ptfem_solve_opts* opts, double* phi, ptfem_solve_stats* stats) {
  PT_ARG(m && phi, "null pointer");
  int rc = ptfem_solve_device(m, opts, stats);
  if (rc != PTFEM_OK && rc != PTFEM_ERR_NOCONV) return rc;
  // hand the (possibly unconverged) iterate back, system-major
  ptfem_ctx* ctx = m->ctx;
  const std::string keep = g_err;
  if (m->S == 1) {
    PT_CK(cudaMemcpyAsync(phi, m->phi.p, (size_t)m->nn * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  } else {
    PT_TRY(m->scratch_d.alloc((size_t)m->nn * m->nsys_user));
    transpose_out_kernel<<<ceil_div(m->nn, 256), 256, 0, ctx->stream>>>(m->phi.p, m->nn, m->S, m->nsys_user, m->scratch_d.p);
    PT_LAUNCH_CHECK(ctx);
    PT_CK(cudaMemcpyAsync(phi, m->scratch_d.p, (size_t)m->nn * m->nsys_user * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  PT_CK(cudaStreamSynchronize(ctx->stream));
  g_err = keep;
  return rc;
}

int ptfem_phi_get_all_async(ptfem_mesh* m, double* phi) {
  PT_ARG(m && phi, "null pointer");
  if (!m->phi.p || m->S < 1) return set_err(PTFEM_ERR_STATE, "no solution on the device");
  ptfem_ctx* ctx = m->ctx;
  PT_CK(cudaSetDevice(ctx->device));
  const double* src = m->phi.p;
  if (m->S != 1) {   // [nn][S] -> [nsys][nn] on the device, then one copy
    PT_TRY(m->scratch_phi.alloc((size_t)m->nn * m->nsys_user));
    transpose_out_kernel<<<ceil_div(m->nn, 256), 256, 0, ctx->stream>>>(m->phi.p, m->nn, m->S, m->nsys_user, m->scratch_phi.p);
    PT_LAUNCH_CHECK(ctx);
    src = m->scratch_phi.p;
  }
  // read-back on the side stream: it overlaps whatever the caller enqueues next (current recovery, metrics, the next mesh)
  PT_CK(cudaEventRecord(ctx->ev_phi_ready, ctx->stream));
  PT_CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_phi_ready, 0));
  PT_CK(cudaMemcpyAsync(phi, src, (size_t)m->nn * m->nsys_user * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream2));
  return PTFEM_OK;
}

int ptfem_phi_get(ptfem_mesh* m, int32_t sys, double* phi) {
  PT_ARG(m && phi, "null pointer");
  if (!m->phi.p) return set_err(PTFEM_ERR_STATE, "no solution on the device");
  PT_ARG(sys >= 0 && sys < m->nsys_user, "system index out of range");
  PT_CK(cudaSetDevice(m->ctx->device));
  DevBuf<double> tmp;
  PT_TRY(tmp.alloc(m->nn));
  gather_sys_kernel<<<ceil_div(m->nn, 256), 256, 0, m->ctx->stream>>>(m->phi.p, m->nn, m->S, sys, tmp.p);
  PT_LAUNCH_CHECK(m->ctx);
  PT_CK(cudaMemcpyAsync(phi, tmp.p, (size_t)m->nn * sizeof(double), cudaMemcpyDeviceToHost, m->ctx->stream));
  PT_CK(cudaStreamSynchronize(m->ctx->stream));
  return PTFEM_OK;
}

int ptfem_phi_set(ptfem_mesh* m, int32_t sys, const double* phi) {
  PT_ARG(m && phi, "null pointer");
  PT_CK(cudaSetDevice(m->ctx->device));
  PT_TRY(prepare_systems(m));
  PT_ARG(sys >= 0 && sys < m->nsys_user, "system index out of range");
  DevBuf<double> tmp;
  PT_TRY(tmp.alloc(m->nn));
  PT_CK(cudaMemcpyAsync(tmp.p, phi, (size_t)m->nn * sizeof(double), cudaMemcpyHostToDevice, m->ctx->stream));
  scatter_sys_kernel<<<ceil_div(m->nn, 256), 256, 0, m->ctx->stream>>>(tmp.p, m->nn, m->S, sys, m->phi.p);
  PT_LAUNCH_CHECK(m->ctx);
  PT_CK(cudaStreamSynchronize(m->ctx->stream));
  m->J_sys = -1;
  m->J_all_valid = false;
  m->J_from_all = false;
  return PTFEM_OK;
}

int ptfem_spmv(ptfem_mesh* m, int32_t sys, int32_t with_bc, int32_t variant, const double* x, double* y) {
  PT_ARG(m && x && y, "null pointer");
  if (m->nval < 1) return set_err(PTFEM_ERR_STATE, "ptfem_assemble has not been called");
  PT_ARG(sys >= 0 && sys < m->nval, "system index out of range");
  ptfem_ctx* ctx = m->ctx;
  PT_CK(cudaSetDevice(ctx->device));
  if (with_bc) PT_TRY(prepare_systems(m));
  // single-system view of the chosen value set
  DevBuf<double> v1, dx, dy;
  const double* vals = with_bc ? m->val_bc.p : m->val_raw.p;
  if (m->nvalp != 1) {
    PT_TRY(v1.alloc(m->nnz + 8));
    PT_CK(cudaMemsetAsync(v1.p + m->nnz, 0, 8 * sizeof(double), ctx->stream));
    gather_sys_kernel<<<ceil_div(m->nnz, 256), 256, 0, ctx->stream>>>(vals, m->nnz, m->nvalp, sys, v1.p);
    PT_LAUNCH_CHECK(ctx);
    vals = v1.p;
  }
  PT_TRY(dx.alloc(m->nn));
  PT_TRY(dy.alloc(m->nn));
  PT_CK(cudaMemcpyAsync(dx.p, x, (size_t)m->nn * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  LinSys A;
  A.nn = m->nn; A.nnz = m->nnz; A.rowptr = m->rowptr.p; A.col = m->col.p; A.val = vals; A.VS = 1; A.S = 1;
  A.stream_rows = m->stream_rows;
  A.stream_cap = m->stream_cap;
  DevBuf<double> pv;
  if (m->has_rowperm) {
    PT_TRY(pv.alloc(m->nnz + 8));
    PT_CK(cudaMemsetAsync(pv.p + m->nnz, 0, 8 * sizeof(double), ctx->stream));
    PT_TRY(permute_values(ctx, m->nn, m->rowptr.p, m->rowid.p, m->prowptr.p, vals, pv.p));
    A.rowid = m->rowid.p;
    A.prowptr = m->prowptr.p;
    A.pcol = m->pcol.p;
    A.pval = pv.p;
  }
  PT_TRY(spmv_launch(ctx, A, variant, dx.p, dy.p, nullptr, false));
  PT_CK(cudaMemcpyAsync(y, dy.p, (size_t)m->nn * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  PT_CK(cudaStreamSynchronize(ctx->stream));
  return PTFEM_OK;
}

int ptfem_spmv_bench(ptfem_mesh* m, int32_t variant, int32_t iters, double* ms_per_launch) {
  PT_ARG(m && ms_per_launch && iters > 0, "bad arguments");
  ptfem_ctx* ctx = m->ctx;
  PT_CK(cudaSetDevice(ctx->device));
  PT_TRY(prepare_systems(m));
  LinSys A;
  make_linsys(m, A);
  PT_TRY(pcg_work_alloc(ctx, m->work, A.nn, A.S, A.VS));
  // x = b (any non-trivial vector), y = work.q
  PT_CK(cudaMemcpyAsync(m->work.p.p, m->b.p, (size_t)m->nn * m->S * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  for (int k = 0; k < 3; ++k) PT_TRY(spmv_launch(ctx, A, variant, m->work.p.p, m->work.q.p, &m->work, false));
  cudaEvent_t e0, e1;
  PT_CK(cudaEventCreate(&e0));
  PT_CK(cudaEventCreate(&e1));
  PT_CK(cudaEventRecord(e0, ctx->stream));
  for (int k = 0; k < iters; ++k) PT_TRY(spmv_launch(ctx, A, variant, m->work.p.p, m->work.q.p, &m->work, false));
  PT_CK(cudaEventRecord(e1, ctx->stream));
  PT_CK(cudaEventSynchronize(e1));
  float ms = 0.f;
  PT_CK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms_per_launch = (double)ms / iters;
  return PTFEM_OK;
}

int ptfem_element_fields(ptfem_mesh* m, int32_t sys, double* E, double* J) {
  PT_ARG(m, "null mesh");
  PT_CK(cudaSetDevice(m->ctx->device));
  PT_TRY(ptfem_do_element_fields(m, sys));
  if (E) PT_CK(cudaMemcpyAsync(E, m->Eelem.p, (size_t)m->nt * 3 * sizeof(double), cudaMemcpyDeviceToHost, m->ctx->stream));
  if (J) PT_CK(cudaMemcpyAsync(J, m->Jelem.p, (size_t)m->nt * 3 * sizeof(double), cudaMemcpyDeviceToHost, m->ctx->stream));
  PT_CK(cudaStreamSynchronize(m->ctx->stream));
  return PTFEM_OK;
}

int ptfem_recover_current(ptfem_mesh* m, int32_t sys, int32_t method, double* J) {
  PT_ARG(m, "null mesh");
  PT_CK(cudaSetDevice(m->ctx->device));
  PT_TRY(ptfem_do_recover(m, sys, method));
  if (J) {
    PT_CK(cudaMemcpyAsync(J, ptfem_current_ptr(m), (size_t)m->nn * 3 * sizeof(double), cudaMemcpyDeviceToHost, m->ctx->stream));
    PT_CK(cudaStreamSynchronize(m->ctx->stream));
  }
  return PTFEM_OK;
}

int ptfem_recover_current_async(ptfem_mesh* m, int32_t sys, int32_t method, double* J) {
  PT_ARG(m && J, "null pointer");
  ptfem_ctx* ctx = m->ctx;
  PT_CK(cudaSetDevice(ctx->device));
  PT_TRY(ptfem_do_recover(m, sys, method));
  // copy on the side stream once the recovery kernels are done; the next recovery waits for it before it
  // overwrites the device buffer, metric kernels of this system only read it and run concurrently
  PT_CK(cudaEventRecord(ctx->ev_j_ready, ctx->stream));
  PT_CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_j_ready, 0));
  PT_CK(cudaMemcpyAsync(J, ptfem_current_ptr(m), (size_t)m->nn * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream2));
  PT_CK(cudaEventRecord(ctx->ev_j_copied, ctx->stream2));
  m->j_copy_pending = true;
  return PTFEM_OK;
}

int ptfem_current_get(ptfem_mesh* m, double* J) {
  PT_ARG(m && J, "null pointer");
  if (!ptfem_current_ptr(m) || m->J_sys < 0) return set_err(PTFEM_ERR_STATE, "no recovered current on the device");
  PT_CK(cudaSetDevice(m->ctx->device));
  PT_CK(cudaMemcpyAsync(J, ptfem_current_ptr(m), (size_t)m->nn * 3 * sizeof(double), cudaMemcpyDeviceToHost, m->ctx->stream));
  PT_CK(cudaStreamSynchronize(m->ctx->stream));
  return PTFEM_OK;
}

int ptfem_recover_current_batch(ptfem_mesh* m, int32_t method, double* J, int32_t wait) {
  PT_ARG(m, "null mesh");
  ptfem_ctx* ctx = m->ctx;
  PT_CK(cudaSetDevice(ctx->device));
  PT_TRY(ptfem_do_recover_batch(m, method));
  if (J) {
    const size_t bytes = (size_t)m->nsys_user * m->nn * 3 * sizeof(double);
    if (wait) {
      PT_CK(cudaMemcpyAsync(J, m->Jall.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
      PT_CK(cudaStreamSynchronize(ctx->stream));
    } else {
      // read-back on the side stream; the caller's next ptfem_ctx_sync completes it, and the next recovery waits for it
      PT_CK(cudaEventRecord(ctx->ev_j_ready, ctx->stream));
      PT_CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_j_ready, 0));
      PT_CK(cudaMemcpyAsync(J, m->Jall.p, bytes, cudaMemcpyDeviceToHost, ctx->stream2));
      PT_CK(cudaEventRecord(ctx->ev_j_copied, ctx->stream2));
      m->j_copy_pending = true;
    }
  }
  return PTFEM_OK;
}

int ptfem_metrics_batch(ptfem_mesh* m, int32_t nreq, const ptfem_metric_req* req, double* out) {
  PT_ARG(m && req && out, "null pointer");
  PT_ARG(nreq >= 1 && nreq <= 256, "1..256 requests per batch");
  PT_CK(cudaSetDevice(m->ctx->device));
  return ptfem_do_metrics_batch(m, nreq, req, out);
}

int ptfem_metric_nodes(ptfem_mesh* m, int32_t sys, int32_t field, double zmin, double zmax, int32_t mode,
                       const ptfem_footprint* fp, int32_t nfp, double scale_r, double out[4]) {
  PT_ARG(m && out, "null pointer");
  PT_ARG(field >= 0 && field <= 3 && mode >= 0 && mode <= 2, "bad field / mode");
  PT_ARG(mode == 0 || (fp && nfp > 0), "footprints required for mode 1/2");
  PT_CK(cudaSetDevice(m->ctx->device));
  return ptfem_do_metric_nodes(m, sys, field, zmin, zmax, mode, fp, mode == 0 ? 0 : nfp, scale_r, out);
}
int ptfem_metric_pad_current(ptfem_mesh* m, int32_t sys, double zmin, const ptfem_footprint* fp, double scale_r, double out[3]) {
  PT_ARG(m && fp && out, "null pointer");
  PT_CK(cudaSetDevice(m->ctx->device));
  return ptfem_do_metric_pad_current(m, sys, zmin, fp, scale_r, out);
}
int ptfem_metric_roi(ptfem_mesh* m, int32_t sys, const double cen[3], double r0, const double* mult, int32_t nmult, double z0,
                     double z1, int32_t include_tris, double* out) {
  PT_ARG(m && cen && mult && out, "null pointer");
  PT_CK(cudaSetDevice(m->ctx->device));
  return ptfem_do_metric_roi(m, sys, cen, r0, mult, nmult, z0, z1, include_tris, out);
}
int ptfem_metric_column_fit(ptfem_mesh* m, int32_t sys, double cx, double cy, double rad, double out[6]) {
  PT_ARG(m && out, "null pointer");
  PT_CK(cudaSetDevice(m->ctx->device));
  return ptfem_do_metric_column_fit(m, sys, cx, cy, rad, out);
}
int ptfem_metric_jstats(ptfem_mesh* m, int32_t sys, double shift, double out[3]) {
  PT_ARG(m && out, "null pointer");
  PT_CK(cudaSetDevice(m->ctx->device));
  return ptfem_do_metric_jstats(m, sys, shift, out);
}
int ptfem_metric_reaction(ptfem_mesh* m, int32_t sys, int32_t bcid, double* current) {
  PT_ARG(m && current, "null pointer");
  PT_CK(cudaSetDevice(m->ctx->device));
  return ptfem_do_metric_reaction(m, sys, bcid, current);
}
int ptfem_sample_polyline(ptfem_mesh* m, int32_t sys, int64_t npts, const double* pts, double* phi_out, double* af_out) {
  PT_ARG(m && (npts == 0 || pts), "null pointer");
  PT_CK(cudaSetDevice(m->ctx->device));
  return ptfem_do_sample_polyline(m, sys, npts, pts, phi_out, af_out);
}

}  // extern "C"
