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
