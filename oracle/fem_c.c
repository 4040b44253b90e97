/*
 * CPU ORACLE, C part (test infrastructure only — never linked into or called by the product path).
 *
 * Plain C + OpenMP restatement of the arithmetic the reference delegates to ElmerSolver
 * (`Procedure = "StatCurrentSolve" "StatCurrentSolver"`, /root/reference/step01_box/case.sif:33-45;
 * driven from step03_ankle_layers/run_layered_sweep.py:1099 and step04_pressure/run_pressure_sweep.py:727):
 *   oc_pattern_*   node-node CSR structure of the P1 stiffness matrix (Elmer's matrix-structure creation)
 *   oc_assemble    K_ij = sum_e sigma_e V_e gradNi.gradNj on linear tets (Elmer type 504), constant
 *                  integrand => quadrature-exact (`Electric Conductivity` per Material, case.sif:56-59)
 *   oc_neumann     `Current Density = g`: b_i += g A_tri / 3        (run_layered_sweep.py:608-611)
 *   oc_dirichlet   `Potential = v`, symmetric elimination            (case.sif:61-71)
 *   oc_pcg         Jacobi-preconditioned CG (stands in for `Linear System Direct Method = UMFPACK`,
 *                  case.sif:41-42: same solution of the SPD system, to the stated residual)
 *   oc_spmv        y = A x
 * It exists because the numpy oracle (fem_oracle.py) cannot assemble the 20 M-tet benchmark mesh in
 * reasonable time/memory; tests/ check it against fem_oracle.py (which is pinned by the reference's
 * analytic step01 case).  PARITY STATUS: the same as fem_oracle.py — pinned by the analytic known
 * answer only; the reference's node-level fields are "parity unpinned" (no mesh/VTU on disk).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may load this.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int32_t* g_n2t_ptr = NULL; /* node -> incident tets (ascending) */
static int32_t* g_n2t = NULL;
static int32_t* g_col = NULL;
static int64_t g_nnz = 0;

int oc_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void oc_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

void oc_free(void) {
  free(g_n2t_ptr); free(g_n2t); free(g_col);
  g_n2t_ptr = g_n2t = g_col = NULL;
  g_nnz = 0;
}

static void sort_i32(int32_t* a, int n) {
  for (int i = 1; i < n; ++i) {
    int32_t v = a[i];
    int j = i - 1;
    while (j >= 0 && a[j] > v) { a[j + 1] = a[j]; --j; }
    a[j + 1] = v;
  }
}

/* Builds the pattern (kept internally), fills rowptr[nn+1], returns nnz (or -1). Columns via oc_pattern_col. */
int64_t oc_pattern_build(int64_t nn, int64_t nt, const int32_t* tets, int32_t* rowptr) {
  oc_free();
  g_n2t_ptr = (int32_t*)calloc((size_t)nn + 1, sizeof(int32_t));
  if (!g_n2t_ptr) return -1;
  for (int64_t k = 0; k < nt * 4; ++k) g_n2t_ptr[tets[k] + 1]++;
  for (int64_t i = 0; i < nn; ++i) g_n2t_ptr[i + 1] += g_n2t_ptr[i];
  g_n2t = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nt * 4 > 0 ? nt * 4 : 1));
  int32_t* cur = (int32_t*)malloc(sizeof(int32_t) * (size_t)nn);
  if (!g_n2t || !cur) return -1;
  memcpy(cur, g_n2t_ptr, sizeof(int32_t) * (size_t)nn);
  for (int64_t e = 0; e < nt; ++e)           /* ascending e => each list is already sorted */
    for (int a = 0; a < 4; ++a) g_n2t[cur[tets[e * 4 + a]]++] = (int32_t)e;
  free(cur);
  /* pass 1: row lengths */
  int32_t* len = (int32_t*)malloc(sizeof(int32_t) * (size_t)nn);
  if (!len) return -1;
#pragma omp parallel
  {
    int cap = 256;
    int32_t* buf = (int32_t*)malloc(sizeof(int32_t) * cap);
#pragma omp for schedule(static)
    for (int64_t i = 0; i < nn; ++i) {
      const int nt_i = g_n2t_ptr[i + 1] - g_n2t_ptr[i];
      if (4 * nt_i + 1 > cap) { cap = 8 * nt_i + 1; buf = (int32_t*)realloc(buf, sizeof(int32_t) * cap); }
      int n = 0;
      buf[n++] = (int32_t)i;
      for (int k = g_n2t_ptr[i]; k < g_n2t_ptr[i + 1]; ++k)
        for (int a = 0; a < 4; ++a) buf[n++] = tets[(int64_t)g_n2t[k] * 4 + a];
      sort_i32(buf, n);
      int u = 1;
      for (int k = 1; k < n; ++k) if (buf[k] != buf[k - 1]) ++u;
      len[i] = u;
    }
    free(buf);
  }
  int64_t nnz = 0;
  for (int64_t i = 0; i < nn; ++i) { rowptr[i] = (int32_t)nnz; nnz += len[i]; }
  rowptr[nn] = (int32_t)nnz;
  free(len);
  if (nnz > 2147483647LL) return -1;
  g_col = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
  if (!g_col) return -1;
#pragma omp parallel
  {
    int cap = 256;
    int32_t* buf = (int32_t*)malloc(sizeof(int32_t) * cap);
#pragma omp for schedule(static)
    for (int64_t i = 0; i < nn; ++i) {
      const int nt_i = g_n2t_ptr[i + 1] - g_n2t_ptr[i];
      if (4 * nt_i + 1 > cap) { cap = 8 * nt_i + 1; buf = (int32_t*)realloc(buf, sizeof(int32_t) * cap); }
      int n = 0;
      buf[n++] = (int32_t)i;
      for (int k = g_n2t_ptr[i]; k < g_n2t_ptr[i + 1]; ++k)
        for (int a = 0; a < 4; ++a) buf[n++] = tets[(int64_t)g_n2t[k] * 4 + a];
      sort_i32(buf, n);
      int32_t* out = g_col + rowptr[i];
      int u = 0;
      out[u++] = buf[0];
      for (int k = 1; k < n; ++k) if (buf[k] != buf[k - 1]) out[u++] = buf[k];
    }
    free(buf);
  }
  g_nnz = nnz;
  return nnz;
}

void oc_pattern_col(int32_t* col) { memcpy(col, g_col, sizeof(int32_t) * (size_t)g_nnz); }

/* shape-function gradients and |volume| of a linear tet */
static double tet_grads(const double* xyz, const int32_t* t, double g[4][3]) {
  const double* p0 = xyz + 3 * (int64_t)t[0]; const double* p1 = xyz + 3 * (int64_t)t[1];
  const double* p2 = xyz + 3 * (int64_t)t[2]; const double* p3 = xyz + 3 * (int64_t)t[3];
  const double a0 = p1[0] - p0[0], a1 = p1[1] - p0[1], a2 = p1[2] - p0[2];
  const double b0 = p2[0] - p0[0], b1 = p2[1] - p0[1], b2 = p2[2] - p0[2];
  const double c0 = p3[0] - p0[0], c1 = p3[1] - p0[1], c2 = p3[2] - p0[2];
  const double bc0 = b1 * c2 - b2 * c1, bc1 = b2 * c0 - b0 * c2, bc2 = b0 * c1 - b1 * c0;
  const double ca0 = c1 * a2 - c2 * a1, ca1 = c2 * a0 - c0 * a2, ca2 = c0 * a1 - c1 * a0;
  const double ab0 = a1 * b2 - a2 * b1, ab1 = a2 * b0 - a0 * b2, ab2 = a0 * b1 - a1 * b0;
  const double det = a0 * bc0 + a1 * bc1 + a2 * bc2;
  const double inv = det != 0.0 ? 1.0 / det : 0.0;
  g[1][0] = bc0 * inv; g[1][1] = bc1 * inv; g[1][2] = bc2 * inv;
  g[2][0] = ca0 * inv; g[2][1] = ca1 * inv; g[2][2] = ca2 * inv;
  g[3][0] = ab0 * inv; g[3][1] = ab1 * inv; g[3][2] = ab2 * inv;
  for (int k = 0; k < 3; ++k) g[0][k] = -(g[1][k] + g[2][k] + g[3][k]);
  return fabs(det) / 6.0;
}

/* val[nnz] on the pattern built by oc_pattern_build; sigma_e[nt] is the per-element conductivity.
 * Row-wise gather in ascending (tet, local index) order: a fixed summation order. Returns 0 / -1. */
int oc_assemble(int64_t nn, int64_t nt, const double* xyz, const int32_t* tets, const double* sigma_e,
                const int32_t* rowptr, double* val) {
  if (!g_col || !g_n2t) return -1;
  (void)nt;
  memset(val, 0, sizeof(double) * (size_t)g_nnz);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nn; ++i) {
    const int32_t rb = rowptr[i], re = rowptr[i + 1];
    for (int k = g_n2t_ptr[i]; k < g_n2t_ptr[i + 1]; ++k) {
      const int64_t e = g_n2t[k];
      const int32_t* t = tets + e * 4;
      double g[4][3];
      const double sv = sigma_e[e] * tet_grads(xyz, t, g);
      for (int a = 0; a < 4; ++a) {
        if (t[a] != (int32_t)i) continue;
        for (int b = 0; b < 4; ++b) {
          int32_t lo = rb, hi = re;
          while (lo < hi) { int32_t mid = (lo + hi) >> 1; if (g_col[mid] < t[b]) lo = mid + 1; else hi = mid; }
          val[lo] += sv * (g[a][0] * g[b][0] + g[a][1] * g[b][1] + g[a][2] * g[b][2]);
        }
      }
    }
  }
  return 0;
}

/* b[i] += g * A / 3 for the nodes of every boundary triangle with bcid == id */
void oc_neumann(const double* xyz, int64_t nb, const int32_t* tris, const int32_t* bcid, int32_t id, double g, double* b) {
  for (int64_t t = 0; t < nb; ++t) {
    if (bcid[t] != id) continue;
    const double* p0 = xyz + 3 * (int64_t)tris[t * 3]; const double* p1 = xyz + 3 * (int64_t)tris[t * 3 + 1];
    const double* p2 = xyz + 3 * (int64_t)tris[t * 3 + 2];
    const double u0 = p1[0] - p0[0], u1 = p1[1] - p0[1], u2 = p1[2] - p0[2];
    const double v0 = p2[0] - p0[0], v1 = p2[1] - p0[1], v2 = p2[2] - p0[2];
    const double c0 = u1 * v2 - u2 * v1, c1 = u2 * v0 - u0 * v2, c2 = u0 * v1 - u1 * v0;
    const double w = g * 0.5 * sqrt(c0 * c0 + c1 * c1 + c2 * c2) / 3.0;
    for (int a = 0; a < 3; ++a) b[tris[t * 3 + a]] += w;
  }
}

/* symmetric elimination in place: b -= K[:,D] v_D ; rows/cols of D zeroed ; diag = 1 ; b_D = v_D */
void oc_dirichlet(int64_t nn, const int32_t* rowptr, const uint8_t* isdir, const double* dirval, double* val, double* b) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nn; ++i) {
    if (isdir[i]) {
      for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) val[k] = g_col[k] == (int32_t)i ? 1.0 : 0.0;
      b[i] = dirval[i];
    } else {
      double acc = b[i];
      for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
        const int32_t j = g_col[k];
        if (isdir[j]) { acc -= val[k] * dirval[j]; val[k] = 0.0; }
      }
      b[i] = acc;
    }
  }
}

void oc_spmv(int64_t nn, const int32_t* rowptr, const int32_t* col, const double* val, const double* x, double* y) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nn; ++i) {
    double acc = 0.0;
    for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) acc += val[k] * x[col[k]];
    y[i] = acc;
  }
}

/* Jacobi-PCG from x (in/out).  Stops at ||r|| <= rtol ||b|| or after maxit iterations.
 * Returns the iteration count; *rel_out = final ||r||/||b|| (recurrence). */
int oc_pcg(int64_t nn, const int32_t* rowptr, const int32_t* col, const double* val, const double* b, double* x,
           double rtol, int maxit, double* rel_out) {
  double* r = (double*)malloc(sizeof(double) * nn);
  double* p = (double*)malloc(sizeof(double) * nn);
  double* q = (double*)malloc(sizeof(double) * nn);
  double* dinv = (double*)malloc(sizeof(double) * nn);
  if (!r || !p || !q || !dinv) return -1;
  double bn2 = 0.0, rz = 0.0, rr = 0.0;
  oc_spmv(nn, rowptr, col, val, x, q);
#pragma omp parallel for schedule(static) reduction(+ : bn2, rz, rr)
  for (int64_t i = 0; i < nn; ++i) {
    double d = 1.0;
    for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) if (col[k] == (int32_t)i) d = val[k];
    dinv[i] = d != 0.0 ? 1.0 / d : 1.0;
    r[i] = b[i] - q[i];
    p[i] = r[i] * dinv[i];
    bn2 += b[i] * b[i];
    rz += r[i] * p[i];
    rr += r[i] * r[i];
  }
  int it = 0;
  while (it < maxit && rr > rtol * rtol * bn2) {
    oc_spmv(nn, rowptr, col, val, p, q);
    double pq = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : pq)
    for (int64_t i = 0; i < nn; ++i) pq += p[i] * q[i];
    const double alpha = pq > 0.0 ? rz / pq : 0.0;
    double rz_new = 0.0;
    rr = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : rz_new, rr)
    for (int64_t i = 0; i < nn; ++i) {
      x[i] += alpha * p[i];
      r[i] -= alpha * q[i];
      rz_new += r[i] * r[i] * dinv[i];
      rr += r[i] * r[i];
    }
    const double beta = rz > 0.0 ? rz_new / rz : 0.0;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nn; ++i) p[i] = r[i] * dinv[i] + beta * p[i];
    rz = rz_new;
    ++it;
  }
  if (rel_out) *rel_out = bn2 > 0.0 ? sqrt(rr / bn2) : 0.0;
  free(r); free(p); free(q); free(dinv);
  return it;
}

/* ------------------------------------------------------------------------------------------------
 * Geometric coarse-grid preconditioner, C/OpenMP restatement of oracle/coarse_oracle.py
 *   M^-1 r = D^-1 r + sum_l Z_l B_l Z_l^T r
 * (Z_l trilinear interpolation from nested regular grids, B of the coarsest grid the exact Galerkin
 * inverse, B_l of the finer grids the inverse Galerkin diagonal; level weights folded into B by the
 * caller).  The reference has no counterpart (it solves directly, step01_box/case.sif:41-42): this only
 * changes how fast PCG converges.  It exists so that bench.py's CPU arm can run the SAME algorithm as
 * the GPU arm on all host cores (like-for-like ratio) and so that the GPU solution on the 20 M-tet mesh
 * can be checked against a CPU solve in seconds.  Nested grids: Z_coarse = Z_fine P, so the mesh is
 * touched once per application (finest grid) and the other levels are reached by grid transfers.
 * ------------------------------------------------------------------------------------------------ */

/* node0[i] = linear index (in the (n+1)^3 node grid) of corner 0 of the cell holding mesh node i,
 * t[i][3] = local coordinates snapped onto the grid planes (coarse_oracle.interpolation) */
void oc_coarse_locate(int64_t nn, const double* xyz, const double* lo, const double* inv_h, const int32_t* n,
                      int32_t* node0, double* t) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nn; ++i) {
    int c[3];
    for (int d = 0; d < 3; ++d) {
      const double u = (xyz[3 * i + d] - lo[d]) * inv_h[d];
      int ci = (int)floor(u);
      ci = ci < 0 ? 0 : (ci > n[d] - 1 ? n[d] - 1 : ci);
      double tt = u - (double)ci;
      tt = tt < 1e-9 ? 0.0 : (tt > 1.0 - 1e-9 ? 1.0 : tt);
      t[3 * i + d] = tt;
      c[d] = ci;
    }
    node0[i] = c[0] + (n[0] + 1) * (c[1] + (n[1] + 1) * c[2]);
  }
}

static inline void corner_weights(const double* t, double w[8]) {
  for (int a = 0; a < 8; ++a)
    w[a] = ((a & 1) ? t[0] : 1.0 - t[0]) * ((a & 2) ? t[1] : 1.0 - t[1]) * ((a & 4) ? t[2] : 1.0 - t[2]);
}
static inline void corner_nodes(int32_t n0, const int32_t* n, int32_t id[8]) {
  const int32_t nx1 = n[0] + 1, nxy = (n[0] + 1) * (n[1] + 1);
  for (int a = 0; a < 8; ++a) id[a] = n0 + (a & 1) + nx1 * ((a >> 1) & 1) + nxy * (a >> 2);
}

/* E[k][k] = Z^T K Z (dense, row-major), k = (n0+1)(n1+1)(n2+1).  Thread-private accumulation, then a
 * reduction in thread order (deterministic for a fixed thread count).  Returns 0 / -1. */
int oc_galerkin_dense(int64_t nn, const int32_t* rowptr, const int32_t* col, const double* val, const uint8_t* isdir,
                      const int32_t* node0, const double* t, const int32_t* n, double* E) {
  const int64_t k = (int64_t)(n[0] + 1) * (n[1] + 1) * (n[2] + 1);
  int nth = 1;
#ifdef _OPENMP
  nth = omp_get_max_threads();
#endif
  double* priv = (double*)calloc((size_t)nth * k * k, sizeof(double));
  if (!priv) return -1;
#pragma omp parallel
  {
    int tid = 0;
#ifdef _OPENMP
    tid = omp_get_thread_num();
#endif
    double* P = priv + (size_t)tid * k * k;
#pragma omp for schedule(static)
    for (int64_t i = 0; i < nn; ++i) {
      if (isdir[i]) continue;
      double wi[8];
      int32_t Ii[8];
      corner_weights(t + 3 * i, wi);
      corner_nodes(node0[i], n, Ii);
      for (int32_t e = rowptr[i]; e < rowptr[i + 1]; ++e) {
        const int32_t j = col[e];
        const double v = val[e];
        if (v == 0.0 || isdir[j]) continue;
        double wj[8];
        int32_t Jj[8];
        corner_weights(t + 3 * (int64_t)j, wj);
        corner_nodes(node0[j], n, Jj);
        for (int a = 0; a < 8; ++a) {
          if (wi[a] == 0.0) continue;
          const double wv = wi[a] * v;
          double* row = P + (size_t)Ii[a] * k;
          for (int b = 0; b < 8; ++b) row[Jj[b]] += wv * wj[b];
        }
      }
    }
  }
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < k * k; ++e) {
    double s = 0.0;
    for (int th = 0; th < nth; ++th) s += priv[(size_t)th * k * k + e];
    E[e] = s;
  }
  free(priv);
  return 0;
}

/* diag[k] of Z^T K Z for a (finer) grid */
int oc_galerkin_diag(int64_t nn, const int32_t* rowptr, const int32_t* col, const double* val, const uint8_t* isdir,
                     const int32_t* node0, const double* t, const int32_t* n, double* diag) {
  const int64_t k = (int64_t)(n[0] + 1) * (n[1] + 1) * (n[2] + 1);
  int nth = 1;
#ifdef _OPENMP
  nth = omp_get_max_threads();
#endif
  double* priv = (double*)calloc((size_t)nth * k, sizeof(double));
  if (!priv) return -1;
#pragma omp parallel
  {
    int tid = 0;
#ifdef _OPENMP
    tid = omp_get_thread_num();
#endif
    double* P = priv + (size_t)tid * k;
#pragma omp for schedule(static)
    for (int64_t i = 0; i < nn; ++i) {
      if (isdir[i]) continue;
      double wi[8];
      int32_t Ii[8];
      corner_weights(t + 3 * i, wi);
      corner_nodes(node0[i], n, Ii);
      for (int32_t e = rowptr[i]; e < rowptr[i + 1]; ++e) {
        const int32_t j = col[e];
        const double v = val[e];
        if (v == 0.0 || isdir[j]) continue;
        double wj[8];
        int32_t Jj[8];
        corner_weights(t + 3 * (int64_t)j, wj);
        corner_nodes(node0[j], n, Jj);
        for (int a = 0; a < 8; ++a)
          for (int b = 0; b < 8; ++b)
            if (Ii[a] == Jj[b]) P[Ii[a]] += wi[a] * v * wj[b];
      }
    }
  }
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < k; ++e) {
    double s = 0.0;
    for (int th = 0; th < nth; ++th) s += priv[(size_t)th * k + e];
    diag[e] = s;
  }
  free(priv);
  return 0;
}

/* grid transfers between nested grids (cells of the finer grid are half the coarser one's):
 * rc[I] = sum_f P[f][I] rf[f]  over the 27 fine nodes around 2I;   ytf[f] = yf[f] + (P ytc)[f] */
static void grid_restrict(const int32_t* nc, const double* rf, double* rc) {
  const int nx1 = nc[0] + 1, ny1 = nc[1] + 1, nz1 = nc[2] + 1;
  const int fx1 = 2 * nc[0] + 1, fy1 = 2 * nc[1] + 1, fz1 = 2 * nc[2] + 1;
#pragma omp parallel for schedule(static) collapse(2)
  for (int iz = 0; iz < nz1; ++iz)
    for (int iy = 0; iy < ny1; ++iy)
      for (int ix = 0; ix < nx1; ++ix) {
        double v = 0.0;
        for (int dz = -1; dz <= 1; ++dz) {
          const int fz = 2 * iz + dz;
          if (fz < 0 || fz >= fz1) continue;
          for (int dy = -1; dy <= 1; ++dy) {
            const int fy = 2 * iy + dy;
            if (fy < 0 || fy >= fy1) continue;
            for (int dx = -1; dx <= 1; ++dx) {
              const int fx = 2 * ix + dx;
              if (fx < 0 || fx >= fx1) continue;
              const double w = (dx ? 0.5 : 1.0) * (dy ? 0.5 : 1.0) * (dz ? 0.5 : 1.0);
              v += w * rf[(size_t)fx + (size_t)fx1 * (fy + (size_t)fy1 * fz)];
            }
          }
        }
        rc[(size_t)ix + (size_t)nx1 * (iy + (size_t)ny1 * iz)] = v;
      }
}
static void grid_prolong_add(const int32_t* nc, const double* yf, const double* ytc, double* ytf) {
  const int nx1 = nc[0] + 1, ny1 = nc[1] + 1;
  const int fx1 = 2 * nc[0] + 1, fy1 = 2 * nc[1] + 1, fz1 = 2 * nc[2] + 1;
#pragma omp parallel for schedule(static) collapse(2)
  for (int fz = 0; fz < fz1; ++fz)
    for (int fy = 0; fy < fy1; ++fy)
      for (int fx = 0; fx < fx1; ++fx) {
        const size_t F = (size_t)fx + (size_t)fx1 * (fy + (size_t)fy1 * fz);
        double v = yf[F];
        const double w = ((fx & 1) ? 0.5 : 1.0) * ((fy & 1) ? 0.5 : 1.0) * ((fz & 1) ? 0.5 : 1.0);
        for (int az = 0; az <= (fz & 1); ++az)
          for (int ay = 0; ay <= (fy & 1); ++ay)
            for (int ax = 0; ax <= (fx & 1); ++ax)
              v += w * ytc[(size_t)(fx / 2 + ax) + (size_t)nx1 * ((fy / 2 + ay) + (size_t)ny1 * (fz / 2 + az))];
        ytf[F] = v;
      }
}

/* PCG preconditioned with Jacobi + coarse grids, from x (in/out).  Levels: 0 = finest ... nlev-1 = coarsest.
 * n[nlev][3] cells per axis (n[l] = 2 n[l+1]); node0/t = position of the mesh nodes in the FINEST grid;
 * bdiag = inverse Galerkin diagonals of levels 0..nlev-2, concatenated; bdense[k][k] = inverse of the coarsest
 * Galerkin matrix (both already multiplied by the level weight).  Stops at ||r|| <= rtol ||b||.
 * Returns the iteration count (or -1); *rel_out = final recurrence ||r|| / ||b||. */
int oc_pcg_coarse(int64_t nn, const int32_t* rowptr, const int32_t* col, const double* val, const double* b, double* x,
                  double rtol, int maxit, double* rel_out, int nlev, const int32_t* n, const int32_t* node0,
                  const double* t, const uint8_t* isdir, const double* bdiag, const double* bdense) {
  if (nlev < 1 || nlev > 6) return -1;
  int nth = 1;
#ifdef _OPENMP
  nth = omp_get_max_threads();
#endif
  int64_t kl[6], off[6];
  int64_t tot = 0;
  for (int l = 0; l < nlev; ++l) {
    kl[l] = (int64_t)(n[3 * l] + 1) * (n[3 * l + 1] + 1) * (n[3 * l + 2] + 1);
    off[l] = tot;
    tot += kl[l];
  }
  const int64_t k0 = kl[0];
  double* r = (double*)malloc(sizeof(double) * nn);
  double* p = (double*)malloc(sizeof(double) * nn);
  double* q = (double*)malloc(sizeof(double) * nn);
  double* z = (double*)malloc(sizeof(double) * nn);
  double* dinv = (double*)malloc(sizeof(double) * nn);
  double* rc = (double*)malloc(sizeof(double) * tot);
  double* yc = (double*)malloc(sizeof(double) * tot);
  double* yt = (double*)malloc(sizeof(double) * tot);
  double* priv = (double*)malloc(sizeof(double) * (size_t)nth * k0);
  if (!r || !p || !q || !z || !dinv || !rc || !yc || !yt || !priv) return -1;
  const int32_t* n0 = n;

#define OC_APPLY_PRECOND()                                                                               \
  do {                                                                                                   \
    _Pragma("omp parallel") {                                                                            \
      int tid = 0;                                                                                       \
      tid = omp_get_thread_num();                                                                        \
      double* P = priv + (size_t)tid * k0;                                                               \
      memset(P, 0, sizeof(double) * k0);                                                                 \
      _Pragma("omp for schedule(static)") for (int64_t i = 0; i < nn; ++i) {                             \
        if (isdir[i]) continue;                                                                          \
        double w[8];                                                                                     \
        int32_t id[8];                                                                                   \
        corner_weights(t + 3 * i, w);                                                                    \
        corner_nodes(node0[i], n0, id);                                                                  \
        const double ri = r[i];                                                                          \
        for (int a = 0; a < 8; ++a) P[id[a]] += w[a] * ri;                                               \
      }                                                                                                  \
    }                                                                                                    \
    _Pragma("omp parallel for schedule(static)") for (int64_t e = 0; e < k0; ++e) {                      \
      double s = 0.0;                                                                                    \
      for (int th = 0; th < nth; ++th) s += priv[(size_t)th * k0 + e];                                   \
      rc[e] = s;                                                                                         \
    }                                                                                                    \
    for (int l = 1; l < nlev; ++l) grid_restrict(n + 3 * l, rc + off[l - 1], rc + off[l]);               \
    for (int l = 0; l < nlev - 1; ++l) {                                                                 \
      _Pragma("omp parallel for schedule(static)") for (int64_t e = 0; e < kl[l]; ++e)                   \
          yc[off[l] + e] = bdiag[off[l] + e] * rc[off[l] + e];                                           \
    }                                                                                                    \
    {                                                                                                    \
      const int64_t kc = kl[nlev - 1];                                                                   \
      const double* rcc = rc + off[nlev - 1];                                                            \
      double* ycc = yc + off[nlev - 1];                                                                  \
      _Pragma("omp parallel for schedule(static)") for (int64_t I = 0; I < kc; ++I) {                    \
        double s = 0.0;                                                                                  \
        const double* row = bdense + (size_t)I * kc;                                                     \
        for (int64_t J = 0; J < kc; ++J) s += row[J] * rcc[J];                                           \
        ycc[I] = s;                                                                                      \
      }                                                                                                  \
      memcpy(yt + off[nlev - 1], ycc, sizeof(double) * kc);                                              \
    }                                                                                                    \
    for (int l = nlev - 2; l >= 0; --l) grid_prolong_add(n + 3 * (l + 1), yc + off[l], yt + off[l + 1], yt + off[l]); \
    _Pragma("omp parallel for schedule(static)") for (int64_t i = 0; i < nn; ++i) {                      \
      double zi = dinv[i] * r[i];                                                                        \
      if (!isdir[i]) {                                                                                   \
        double w[8];                                                                                     \
        int32_t id[8];                                                                                   \
        corner_weights(t + 3 * i, w);                                                                    \
        corner_nodes(node0[i], n0, id);                                                                  \
        for (int a = 0; a < 8; ++a) zi += w[a] * yt[id[a]];                                              \
      }                                                                                                  \
      z[i] = zi;                                                                                         \
    }                                                                                                    \
  } while (0)

  double bn2 = 0.0, rr = 0.0;
  oc_spmv(nn, rowptr, col, val, x, q);
#pragma omp parallel for schedule(static) reduction(+ : bn2, rr)
  for (int64_t i = 0; i < nn; ++i) {
    double d = 1.0;
    for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) if (col[k] == (int32_t)i) d = val[k];
    dinv[i] = d != 0.0 ? 1.0 / d : 1.0;
    r[i] = b[i] - q[i];
    bn2 += b[i] * b[i];
    rr += r[i] * r[i];
  }
  OC_APPLY_PRECOND();
  double rz = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : rz)
  for (int64_t i = 0; i < nn; ++i) {
    p[i] = z[i];
    rz += r[i] * z[i];
  }
  int it = 0;
  while (it < maxit && rr > rtol * rtol * bn2) {
    oc_spmv(nn, rowptr, col, val, p, q);
    double pq = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : pq)
    for (int64_t i = 0; i < nn; ++i) pq += p[i] * q[i];
    const double alpha = pq > 0.0 ? rz / pq : 0.0;
    rr = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : rr)
    for (int64_t i = 0; i < nn; ++i) {
      x[i] += alpha * p[i];
      r[i] -= alpha * q[i];
      rr += r[i] * r[i];
    }
    OC_APPLY_PRECOND();
    double rz_new = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : rz_new)
    for (int64_t i = 0; i < nn; ++i) rz_new += r[i] * z[i];
    const double beta = rz > 0.0 ? rz_new / rz : 0.0;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nn; ++i) p[i] = z[i] + beta * p[i];
    rz = rz_new;
    ++it;
  }
#undef OC_APPLY_PRECOND
  if (rel_out) *rel_out = bn2 > 0.0 ? sqrt(rr / bn2) : 0.0;
  free(r); free(p); free(q); free(z); free(dinv); free(rc); free(yc); free(yt); free(priv);
  return it;
}

/* `Calculate Volume Current = True` (step01_box/case.sif:39), volume-weighted ("lumped") nodal recovery as
 * fem_oracle.recover_nodal_current(method="lumped"): J_i = sum_e (V_e/4) J_e / sum_e (V_e/4), J_e = -sigma_e grad phi.
 * Uses the node -> tet lists of the last oc_pattern_build (same mesh).  Returns 0 / -1. */
int oc_recover_lumped(int64_t nn, const double* xyz, const int32_t* tets, const double* sigma_e, const double* phi,
                      double* J) {
  if (!g_n2t || !g_n2t_ptr) return -1;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nn; ++i) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, m = 0.0;
    for (int k = g_n2t_ptr[i]; k < g_n2t_ptr[i + 1]; ++k) {
      const int64_t e = g_n2t[k];
      const int32_t* tt = tets + e * 4;
      double g[4][3];
      const double vol = tet_grads(xyz, tt, g);
      double ex = 0.0, ey = 0.0, ez = 0.0;
      for (int a = 0; a < 4; ++a) {
        const double v = phi[tt[a]];
        ex -= v * g[a][0]; ey -= v * g[a][1]; ez -= v * g[a][2];
      }
      const double w = 0.25 * vol, sg = sigma_e[e];
      a0 += w * sg * ex; a1 += w * sg * ey; a2 += w * sg * ez;
      m += w;
    }
    const double d = m > 0.0 ? 1.0 / m : 0.0;
    J[3 * i] = a0 * d; J[3 * i + 1] = a1 * d; J[3 * i + 2] = a2 * d;
  }
  return 0;
}
