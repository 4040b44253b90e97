// assembly.cu — K2 element geometry factors, K3 gather-form stiffness assembly (no atomics,
// fixed summation order), K4 boundary conditions.
//
// Replaces the bulk / boundary assembly of Elmer's StatCurrentSolver (weak form of
// div(sigma grad phi) = 0 on linear tets, `Electric Conductivity` per Material:
// step03_ankle_layers/run_layered_sweep.py:563-587), the Neumann `Current Density` load
// (:608-611) and the Dirichlet `Potential` condition (:620-624).
#include "common.cuh"
#include "geom.cuh"

using namespace ptfem;

namespace {

// K2: per tet |V| and G[ij] = |V| gradNi.gradNj (10 unique)
__global__ void geom_kernel(const double* __restrict__ xyz, const int32_t* __restrict__ tets, int64_t nt,
                            double* __restrict__ G, double* __restrict__ vol, int32_t* __restrict__ nbad,
                            unsigned long long* __restrict__ hmax2) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // longest edge of the mesh (squared, as ordered bits of a non-negative double): bounds how far a node can be
  // from the centroid of a cell that uses it (lazy smoothing of the ROI metric)
  double l2 = 0.0;
  if (e < nt) {
    const int32_t* t = tets + e * 4;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = a + 1; b < 4; ++b) {
        const double dx = xyz[3 * (int64_t)t[a]] - xyz[3 * (int64_t)t[b]], dy = xyz[3 * (int64_t)t[a] + 1] - xyz[3 * (int64_t)t[b] + 1],
                     dz = xyz[3 * (int64_t)t[a] + 2] - xyz[3 * (int64_t)t[b] + 2];
        l2 = fmax(l2, dx * dx + dy * dy + dz * dz);
      }
  }
  l2 = warp_max(l2);
  if ((threadIdx.x & 31) == 0 && l2 > 0.0) atomicMax(hmax2, (unsigned long long)__double_as_longlong(l2));
  if (e >= nt) return;
  double g[4][3];
  const double v = fabs(tet_grads(xyz, tets + e * 4, g));
  if (!(v > 0.0)) atomicAdd(nbad, 1);
  vol[e] = v;
  int idx = 0;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = a; b < 4; ++b) {
      G[e * 10 + idx] = v * (g[a][0] * g[b][0] + g[a][1] * g[b][1] + g[a][2] * g[b][2]);
      ++idx;
    }
}

__global__ void tri_area_kernel(const double* __restrict__ xyz, const int32_t* __restrict__ tris, int64_t nb,
                                double* __restrict__ area) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nb) return;
  const int64_t n0 = tris[t * 3], n1 = tris[t * 3 + 1], n2 = tris[t * 3 + 2];
  const double u0 = xyz[n1 * 3] - xyz[n0 * 3], u1 = xyz[n1 * 3 + 1] - xyz[n0 * 3 + 1], u2 = xyz[n1 * 3 + 2] - xyz[n0 * 3 + 2];
  const double v0 = xyz[n2 * 3] - xyz[n0 * 3], v1 = xyz[n2 * 3 + 1] - xyz[n0 * 3 + 1], v2 = xyz[n2 * 3 + 2] - xyz[n0 * 3 + 2];
  const double c0 = u1 * v2 - u2 * v1, c1 = u2 * v0 - u0 * v2, c2 = u0 * v1 - u1 * v0;
  area[t] = 0.5 * sqrt(c0 * c0 + c1 * c1 + c2 * c2);
}

// lumped mass (sum of V/4 over incident tets, ascending tet order) and valence, thread per node
__global__ void node_mass_kernel(const int32_t* __restrict__ n2t_ptr, const int32_t* __restrict__ n2t,
                                 const double* __restrict__ vol, int64_t nn, double* __restrict__ mlump,
                                 int32_t* __restrict__ valence) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  const int32_t b = n2t_ptr[i], e = n2t_ptr[i + 1];
  double s = 0.0;
  for (int32_t k = b; k < e; ++k) s += 0.25 * vol[n2t[k]];
  mlump[i] = s;
  valence[i] = e - b;
}

__global__ void region_index_kernel(const int32_t* __restrict__ region, int64_t nt, const int32_t* __restrict__ reg_ids,
                                    int nreg, uint8_t* __restrict__ regidx, int32_t* __restrict__ nmissing) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nt) return;
  const int32_t r = region[e];
  int found = -1;
  for (int k = 0; k < nreg; ++k)
    if (reg_ids[k] == r) found = k;
  if (found < 0) {
    atomicAdd(nmissing, 1);
    found = 0;
  }
  regidx[e] = (uint8_t)found;
}

// K3: val[k][s] = sum over contributions (tet, ij) of sigma_s[region[tet]] * G[tet][sym(ij)], ascending (tet, ij).
template <int S>
__global__ void assemble_kernel(const int32_t* __restrict__ gptr, const int32_t* __restrict__ gsrc,
                                const double* __restrict__ G, const uint8_t* __restrict__ regidx,
                                const double* __restrict__ sigma_tab, int nreg, int64_t nnz, double* __restrict__ val) {
  extern __shared__ double s_sig[];  // [S][nreg]
  for (int i = threadIdx.x; i < S * nreg; i += blockDim.x) s_sig[i] = sigma_tab[i];
  __syncthreads();
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  double acc[S];
#pragma unroll
  for (int s = 0; s < S; ++s) acc[s] = 0.0;
  const int32_t b = gptr[k], e = gptr[k + 1];
  for (int32_t c = b; c < e; ++c) {
    const int32_t src = gsrc[c];
    const int64_t tet = src >> 4;
    const int ij = src & 15;
    const double gv = G[tet * 10 + sym10(ij >> 2, ij & 3)];
    const int r = regidx[tet];
#pragma unroll
    for (int s = 0; s < S; ++s) acc[s] += s_sig[s * nreg + r] * gv;
  }
#pragma unroll
  for (int s = 0; s < S; ++s) val[k * S + s] = acc[s];
}

// consistent mass values on the same pattern: M_ij = sum_e V_e (1 + delta_ij)/20
__global__ void assemble_mass_kernel(const int32_t* __restrict__ gptr, const int32_t* __restrict__ gsrc,
                                     const double* __restrict__ vol, int64_t nnz, double* __restrict__ val) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  double acc = 0.0;
  const int32_t b = gptr[k], e = gptr[k + 1];
  for (int32_t c = b; c < e; ++c) {
    const int32_t src = gsrc[c];
    const int ij = src & 15;
    acc += vol[src >> 4] * (((ij >> 2) == (ij & 3)) ? 0.1 : 0.05);
  }
  val[k] = acc;
}

// ---- K4 ------------------------------------------------------------------------------------------
__global__ void mark_dirichlet_kernel(const int32_t* __restrict__ tris, const int32_t* __restrict__ bcid, int64_t nb,
                                      int32_t id, double value, int rhs, int S, uint8_t* __restrict__ isdir,
                                      double* __restrict__ dirval) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nb || bcid[t] != id) return;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int64_t n = tris[t * 3 + a];
    isdir[n] = 1;
    if (rhs < 0) {
      for (int s = 0; s < S; ++s) dirval[n * S + s] = value;
    } else {
      dirval[n * S + rhs] = value;
    }
  }
}

__global__ void set_tri_load_bcid(const int32_t* __restrict__ bcid, int64_t nb, int32_t id, double g, int rhs, int S,
                                  double* __restrict__ tri_load) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nb || bcid[t] != id) return;
  if (rhs < 0) {
    for (int s = 0; s < S; ++s) tri_load[(int64_t)s * nb + t] = g;
  } else {
    tri_load[(int64_t)rhs * nb + t] = g;
  }
}

__global__ void set_tri_load_list(const int32_t* __restrict__ idx, int64_t n, int64_t nb, double g, int rhs, int S,
                                  double* __restrict__ tri_load) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t t = idx[i];
  if (t < 0 || t >= nb) return;
  if (rhs < 0) {
    for (int s = 0; s < S; ++s) tri_load[(int64_t)s * nb + t] = g;
  } else {
    tri_load[(int64_t)rhs * nb + t] = g;
  }
}

__global__ void any_dirichlet_kernel(const uint8_t* __restrict__ isdir, int64_t nn, int32_t* __restrict__ flag) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int has = i < nn && isdir[i];
  if (__any_sync(0xffffffffu, has) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

// Neumann load vector, gather over the node's boundary triangles in ascending triangle order.
__global__ void neumann_load_kernel(const int32_t* __restrict__ n2b_ptr, const int32_t* __restrict__ n2b,
                                    const double* __restrict__ tri_area, const double* __restrict__ tri_load,
                                    int64_t nn, int64_t nb, int S, double* __restrict__ b_neu) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  const int32_t b = n2b_ptr[i], e = n2b_ptr[i + 1];
  for (int s = 0; s < S; ++s) {
    double acc = 0.0;
    for (int32_t k = b; k < e; ++k) {
      const int64_t t = n2b[k];
      acc += tri_load[(int64_t)s * nb + t] * tri_area[t] * (1.0 / 3.0);
    }
    b_neu[i * S + s] = acc;
  }
}

// Symmetric Dirichlet elimination, one thread per row:
//   b_i = b_neu_i - sum_{j in D} K_ij phi_j ;  K_iD = K_Di = 0 ;  K_DD = I ;  b_D = phi_D.
// VS = value sets (1: shared matrix, multi-RHS; else == S), S = systems.
__global__ void eliminate_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                 const int32_t* __restrict__ diag, const uint8_t* __restrict__ isdir,
                                 const double* __restrict__ dirval, int dirS, const double* __restrict__ b_neu,
                                 const double* __restrict__ val_raw, int VS, int S, int64_t nn,
                                 double* __restrict__ val_bc, double* __restrict__ b, double* __restrict__ dinv) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  const int32_t rb = rowptr[i], re = rowptr[i + 1], dg = diag[i];
  const bool di = isdir[i] != 0;
  if (VS == 1 && !di) {
    // one shared matrix (multi-RHS sweep): the row is walked ONCE - values, diagonal and whether any neighbour is a
    // Dirichlet node; only rows next to an electrode (a fraction of a percent) then visit their Dirichlet neighbours per system
    double d = 0.0;
    bool any = false;
    for (int32_t k = rb; k < re; ++k) {
      const double v = val_raw[k];
      const bool dj = isdir[col[k]] != 0;
      any = any || dj;
      val_bc[k] = dj ? 0.0 : v;
      if (k == dg && !dj) d = v;
    }
    if (d == 0.0) {  // isolated node (no tets): identity row
      d = 1.0;
      val_bc[dg] = 1.0;
    }
    dinv[i] = 1.0 / d;
    for (int s = 0; s < S; ++s) {
      const int ds = (s < dirS) ? s : 0;
      double acc = b_neu[i * dirS + ds];
      if (any)
        for (int32_t k = rb; k < re; ++k) {
          const int32_t j = col[k];
          if (isdir[j]) acc -= val_raw[k] * dirval[(int64_t)j * dirS + ds];
        }
      b[i * S + s] = acc;
    }
    return;
  }
  for (int s = 0; s < S; ++s) {
    const int vs = (VS == 1) ? 0 : s;
    const int ds = (s < dirS) ? s : 0;
    const bool write_val = (VS != 1) || (s == 0);
    if (di) {
      if (write_val)
        for (int32_t k = rb; k < re; ++k) val_bc[(int64_t)k * VS + vs] = (k == dg) ? 1.0 : 0.0;
      b[i * S + s] = dirval[i * dirS + ds];
      if (write_val) dinv[i * VS + vs] = 1.0;
    } else {
      double acc = b_neu[i * dirS + ds];
      double d = 0.0;
      for (int32_t k = rb; k < re; ++k) {
        const int32_t j = col[k];
        const double v = val_raw[(int64_t)k * VS + vs];
        if (isdir[j]) {
          acc -= v * dirval[(int64_t)j * dirS + ds];
          if (write_val) val_bc[(int64_t)k * VS + vs] = 0.0;
        } else {
          if (write_val) val_bc[(int64_t)k * VS + vs] = v;
          if (k == dg) d = v;
        }
      }
      if (d == 0.0) {  // isolated node (no tets): identity row
        d = 1.0;
        if (write_val) val_bc[(int64_t)dg * VS + vs] = 1.0;
      }
      b[i * S + s] = acc;
      if (write_val) dinv[i * VS + vs] = 1.0 / d;
    }
  }
}

}  // namespace

// -------------------------------------------------------------------------------------------------
int ptfem_build_geometry(ptfem_mesh* m) {
  ptfem_ctx* ctx = m->ctx;
  PT_TRY(m->G.alloc((size_t)m->nt * 10));
  PT_TRY(m->vol.alloc(m->nt));
  PT_TRY(m->tri_area.alloc(m->nb));
  PT_TRY(m->mlump.alloc(m->nn));
  PT_TRY(m->valence.alloc(m->nn));
  DevBuf<int32_t> nbad;
  PT_TRY(nbad.alloc(1));
  PT_TRY(fill_i32(ctx, nbad.p, 0, 1));
  DevBuf<unsigned long long> hmax2;
  PT_TRY(hmax2.alloc(1));
  PT_CK(cudaMemsetAsync(hmax2.p, 0, sizeof(unsigned long long), ctx->stream));
  if (m->nt > 0) {
    geom_kernel<<<ceil_div(m->nt, 128), 128, 0, ctx->stream>>>(m->xyz.p, m->tets.p, m->nt, m->G.p, m->vol.p, nbad.p, hmax2.p);
    PT_LAUNCH_CHECK(ctx);
  }
  if (m->nb > 0) {
    tri_area_kernel<<<ceil_div(m->nb, 128), 128, 0, ctx->stream>>>(m->xyz.p, m->tris.p, m->nb, m->tri_area.p);
    PT_LAUNCH_CHECK(ctx);
  }
  node_mass_kernel<<<ceil_div(m->nn, 128), 128, 0, ctx->stream>>>(m->n2t_ptr.p, m->n2t.p, m->vol.p, m->nn, m->mlump.p,
                                                                   m->valence.p);
  PT_LAUNCH_CHECK(ctx);
  int32_t bad = 0;
  unsigned long long hbits = 0;
  PT_CK(cudaMemcpyAsync(&bad, nbad.p, sizeof bad, cudaMemcpyDeviceToHost, ctx->stream));
  PT_CK(cudaMemcpyAsync(&hbits, hmax2.p, sizeof hbits, cudaMemcpyDeviceToHost, ctx->stream));
  PT_CK(cudaStreamSynchronize(ctx->stream));
  {
    double l2;
    memcpy(&l2, &hbits, sizeof l2);
    m->h_max = sqrt(l2);
  }
  if (bad > 0) return set_err(PTFEM_ERR_ARG, "mesh has %d degenerate (zero-volume) tetrahedra", bad);
  m->has_geom = true;
  m->mval.release();
  return PTFEM_OK;
}

template <int S>
static int launch_assemble(ptfem_mesh* m) {
  ptfem_ctx* ctx = m->ctx;
  assemble_kernel<S><<<ceil_div(m->nnz, 128), 128, S * m->nreg * sizeof(double), ctx->stream>>>(
      m->gptr.p, m->gsrc.p, m->G.p, m->regidx.p, m->sigma_tab.p, m->nreg, m->nnz, m->val_raw.p);
  PT_LAUNCH_CHECK(ctx);
  return PTFEM_OK;
}

int ptfem_do_assemble(ptfem_mesh* m, int32_t nreg, const int32_t* reg_ids, const double* sigma, int32_t nsys) {
  ptfem_ctx* ctx = m->ctx;
  const int Sp = ptfem_pad_nsys(nsys);
  if (Sp < 0) return set_err(PTFEM_ERR_ARG, "at most 16 systems per batch (got %d)", nsys);
  if (nreg < 1 || nreg > 255) return set_err(PTFEM_ERR_ARG, "nreg must be in 1..255");
  DevBuf<int32_t> d_ids, d_missing;
  PT_TRY(d_ids.alloc(nreg));
  PT_TRY(d_missing.alloc(1));
  PT_CK(cudaMemcpyAsync(d_ids.p, reg_ids, nreg * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  PT_TRY(fill_i32(ctx, d_missing.p, 0, 1));
  PT_TRY(m->regidx.alloc(m->nt));
  if (m->nt > 0) {
    region_index_kernel<<<ceil_div(m->nt, 256), 256, 0, ctx->stream>>>(m->region.p, m->nt, d_ids.p, nreg, m->regidx.p,
                                                                       d_missing.p);
    PT_LAUNCH_CHECK(ctx);
  }
  int32_t missing = 0;
  PT_CK(cudaMemcpyAsync(&missing, d_missing.p, sizeof missing, cudaMemcpyDeviceToHost, ctx->stream));
  PT_CK(cudaStreamSynchronize(ctx->stream));
  if (missing > 0)
    return set_err(PTFEM_ERR_ARG, "%d elements belong to a body without a conductivity (SIF Body/Material)", missing);
  // sigma table padded: extra systems repeat system 0 (keeps them SPD; their rhs is zero)
  std::vector<double> tab((size_t)Sp * nreg);
  for (int s = 0; s < Sp; ++s)
    for (int r = 0; r < nreg; ++r) {
      const double v = sigma[(size_t)(s < nsys ? s : 0) * nreg + r];
      if (!(v > 0.0)) return set_err(PTFEM_ERR_ARG, "conductivity must be positive (system %d, region %d)", s, reg_ids[r]);
      tab[(size_t)s * nreg + r] = v;
    }
  PT_TRY(m->sigma_tab.alloc(tab.size()));
  PT_CK(cudaMemcpyAsync(m->sigma_tab.p, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  PT_CK(cudaStreamSynchronize(ctx->stream));
  m->nreg = nreg;
  m->nval = nsys;
  m->nvalp = Sp;
  PT_TRY(m->val_raw.alloc((size_t)(m->nnz + 8) * Sp));
  PT_TRY(m->val_bc.alloc((size_t)(m->nnz + 8) * Sp));
  PT_CK(cudaMemsetAsync(m->val_raw.p + (size_t)m->nnz * Sp, 0, (size_t)8 * Sp * sizeof(double), ctx->stream));
  PT_CK(cudaMemsetAsync(m->val_bc.p + (size_t)m->nnz * Sp, 0, (size_t)8 * Sp * sizeof(double), ctx->stream));
  switch (Sp) {
    case 1: PT_TRY(launch_assemble<1>(m)); break;
    case 2: PT_TRY(launch_assemble<2>(m)); break;
    case 4: PT_TRY(launch_assemble<4>(m)); break;
    case 8: PT_TRY(launch_assemble<8>(m)); break;
    default: PT_TRY(launch_assemble<16>(m)); break;
  }
  m->bc_dirty = true;
  return PTFEM_OK;
}

int ptfem_do_assemble_mass(ptfem_mesh* m) {
  if (m->mval.p && m->mval.n >= (size_t)m->nnz) return PTFEM_OK;
  ptfem_ctx* ctx = m->ctx;
  PT_TRY(m->mval.alloc(m->nnz + 8));
  PT_CK(cudaMemsetAsync(m->mval.p + m->nnz, 0, 8 * sizeof(double), ctx->stream));
  assemble_mass_kernel<<<ceil_div(m->nnz, 128), 128, 0, ctx->stream>>>(m->gptr.p, m->gsrc.p, m->vol.p, m->nnz, m->mval.p);
  PT_LAUNCH_CHECK(ctx);
  return PTFEM_OK;
}

int ptfem_do_bc_reset(ptfem_mesh* m, int32_t nrhs) {
  ptfem_ctx* ctx = m->ctx;
  const int Sp = ptfem_pad_nsys(nrhs);
  if (Sp < 0) return set_err(PTFEM_ERR_ARG, "at most 16 right-hand sides per batch (got %d)", nrhs);
  m->nrhs = nrhs;
  m->nrhsp = Sp;
  PT_TRY(m->isdir.alloc(m->nn));
  PT_TRY(m->dirval.alloc((size_t)m->nn * Sp));
  PT_TRY(m->tri_load.alloc((size_t)(m->nb > 0 ? m->nb : 1) * Sp));
  PT_TRY(m->b_neu.alloc((size_t)m->nn * Sp));
  PT_CK(cudaMemsetAsync(m->isdir.p, 0, m->nn, ctx->stream));
  PT_CK(cudaMemsetAsync(m->dirval.p, 0, (size_t)m->nn * Sp * sizeof(double), ctx->stream));
  PT_CK(cudaMemsetAsync(m->tri_load.p, 0, (size_t)(m->nb > 0 ? m->nb : 1) * Sp * sizeof(double), ctx->stream));
  m->bc_dirty = true;
  return PTFEM_OK;
}

int ptfem_do_bc_dirichlet(ptfem_mesh* m, int32_t rhs, int32_t bcid, double value) {
  ptfem_ctx* ctx = m->ctx;
  if (m->nb > 0) {
    mark_dirichlet_kernel<<<ceil_div(m->nb, 256), 256, 0, ctx->stream>>>(m->tris.p, m->bcid.p, m->nb, bcid, value, rhs,
                                                                         m->nrhsp, m->isdir.p, m->dirval.p);
    PT_LAUNCH_CHECK(ctx);
  }
  m->bc_dirty = true;
  return PTFEM_OK;
}

int ptfem_do_bc_neumann(ptfem_mesh* m, int32_t rhs, int32_t bcid, double g) {
  ptfem_ctx* ctx = m->ctx;
  if (m->nb > 0) {
    set_tri_load_bcid<<<ceil_div(m->nb, 256), 256, 0, ctx->stream>>>(m->bcid.p, m->nb, bcid, g, rhs, m->nrhsp,
                                                                     m->tri_load.p);
    PT_LAUNCH_CHECK(ctx);
  }
  m->bc_dirty = true;
  return PTFEM_OK;
}

int ptfem_do_bc_neumann_tris(ptfem_mesh* m, int32_t rhs, int64_t n, const int32_t* tri_idx, double g) {
  ptfem_ctx* ctx = m->ctx;
  if (n <= 0) return PTFEM_OK;
  DevBuf<int32_t> d;
  PT_TRY(d.alloc(n));
  PT_CK(cudaMemcpyAsync(d.p, tri_idx, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  set_tri_load_list<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(d.p, n, m->nb, g, rhs, m->nrhsp, m->tri_load.p);
  PT_LAUNCH_CHECK(ctx);
  PT_CK(cudaStreamSynchronize(ctx->stream));
  m->bc_dirty = true;
  return PTFEM_OK;
}

// Builds b (and the eliminated matrix) for S = max(nvalp, nrhsp) systems.
int ptfem_apply_bc(ptfem_mesh* m, double* dinv_out /*[nn][VS]*/, int* S_out) {
  ptfem_ctx* ctx = m->ctx;
  if (m->nval < 1) return set_err(PTFEM_ERR_STATE, "ptfem_assemble has not been called");
  if (m->nrhs < 1) return set_err(PTFEM_ERR_STATE, "ptfem_bc_reset has not been called");
  const int VS = m->nvalp;
  const int S = VS > m->nrhsp ? VS : m->nrhsp;
  if (VS != 1 && m->nrhsp != 1 && VS != m->nrhsp)
    return set_err(PTFEM_ERR_ARG, "batched matrices (%d) and right-hand sides (%d) must match or one of them be 1", m->nval,
                   m->nrhs);
  {  // a problem without any `Potential` node is singular (pure Neumann): refuse instead of iterating forever
    DevBuf<int32_t> flag;
    PT_TRY(flag.alloc(1));
    PT_TRY(fill_i32(ctx, flag.p, 0, 1));
    any_dirichlet_kernel<<<ceil_div(m->nn, 256), 256, 0, ctx->stream>>>(m->isdir.p, m->nn, flag.p);
    PT_LAUNCH_CHECK(ctx);
    int32_t h = 0;
    PT_CK(cudaMemcpyAsync(&h, flag.p, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    PT_CK(cudaStreamSynchronize(ctx->stream));
    if (!h && !m->is_part)   // (one rank's part of a larger mesh may hold no electrode at all)
      return set_err(PTFEM_ERR_ARG, "no mesh node carries a Potential (Dirichlet) condition: the boundary ids given to "
                                    "ptfem_bc_dirichlet do not occur in the mesh, or none was set (pure Neumann problem is singular)");
  }
  neumann_load_kernel<<<ceil_div(m->nn, 128), 128, 0, ctx->stream>>>(m->n2b_ptr.p, m->n2b.p, m->tri_area.p, m->tri_load.p,
                                                                      m->nn, m->nb, m->nrhsp, m->b_neu.p);
  PT_LAUNCH_CHECK(ctx);
  PT_TRY(m->b.alloc((size_t)m->nn * S));
  eliminate_kernel<<<ceil_div(m->nn, 128), 128, 0, ctx->stream>>>(m->rowptr.p, m->col.p, m->diag.p, m->isdir.p, m->dirval.p,
                                                                   m->nrhsp, m->b_neu.p, m->val_raw.p, VS, S, m->nn,
                                                                   m->val_bc.p, m->b.p, dinv_out);
  PT_LAUNCH_CHECK(ctx);
  *S_out = S;
  m->bc_dirty = false;
  m->matrix_epoch++;
  return PTFEM_OK;
}
