// post.cu — K10 element fields, K11 nodal current recovery, K12 metric reductions, K13 polyline
// sampling / activating function.
//
// Replaces (a) Elmer's `Calculate Volume Current = True` (step01_box/case.sif:39): nodal J = -sigma grad phi,
// and (b) the pyvista/VTK filters the reference's metric extraction runs on the VTU
// (step03_ankle_layers/run_layered_sweep.py:704-1030, step04_pressure/run_pressure_sweep.py:446-660,
// step02_electrodes/run_sweep.py:286-295, step01_box/test_step01_baseline.py:59-104).
// All gathers walk the sorted node->element lists, so sums have a fixed order (bit-reproducible).
#include <cmath>

#include "geom.cuh"
#include "solver.cuh"

using namespace ptfem;

namespace {

constexpr int kT = 256;

// ---- K10 -------------------------------------------------------------------------------------------
__global__ void element_fields_kernel(const double* __restrict__ xyz, const int32_t* __restrict__ tets, int64_t nt,
                                      const double* __restrict__ phi, int S, int sys, const uint8_t* __restrict__ regidx,
                                      const double* __restrict__ sigma, double* __restrict__ E, double* __restrict__ J) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nt) return;
  double g[4][3];
  tet_grads(xyz, tets + e * 4, g);
  double ex = 0.0, ey = 0.0, ez = 0.0;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const double v = phi[(int64_t)tets[e * 4 + a] * S + sys];
    ex -= v * g[a][0];
    ey -= v * g[a][1];
    ez -= v * g[a][2];
  }
  const double sg = sigma[regidx[e]];
  E[e * 3 + 0] = ex; E[e * 3 + 1] = ey; E[e * 3 + 2] = ez;
  J[e * 3 + 0] = sg * ex; J[e * 3 + 1] = sg * ey; J[e * 3 + 2] = sg * ez;
}

// ---- K11 -------------------------------------------------------------------------------------------
// mode 0: rhs4[i][k] = sum_e (V_e/4) J_e^k (k<3), [3] = 0     (L2 right-hand side, padded to 4 systems)
// mode 1: out3[i][k] = that / mlump_i                         (lumped)
// mode 2: out3[i][k] = sum_e J_e^k / valence_i                (unweighted average)
__global__ void recover_gather_kernel(const int32_t* __restrict__ n2t_ptr, const int32_t* __restrict__ n2t,
                                      const double* __restrict__ vol, const double* __restrict__ Je,
                                      const double* __restrict__ mlump, int64_t nn, int mode, double* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  const int32_t b = n2t_ptr[i], e = n2t_ptr[i + 1];
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  for (int32_t k = b; k < e; ++k) {
    const int64_t t = n2t[k];
    const double w = mode == 2 ? 1.0 : 0.25 * vol[t];
    a0 += w * Je[t * 3 + 0];
    a1 += w * Je[t * 3 + 1];
    a2 += w * Je[t * 3 + 2];
  }
  if (mode == 0) {
    out[i * 4 + 0] = a0; out[i * 4 + 1] = a1; out[i * 4 + 2] = a2; out[i * 4 + 3] = 0.0;
  } else {
    double d = mode == 1 ? mlump[i] : (double)(e - b);
    d = d > 0.0 ? 1.0 / d : 0.0;
    out[i * 3 + 0] = a0 * d; out[i * 3 + 1] = a1 * d; out[i * 3 + 2] = a2 * d;
  }
}

// inverse diagonal of the mass matrix; nodes without tets get an identity row
__global__ void mass_dinv_kernel(const int32_t* __restrict__ diag, double* __restrict__ mval, int64_t nn,
                                 double* __restrict__ dinv) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  double d = mval[diag[i]];
  if (!(d > 0.0)) {
    d = 1.0;
    mval[diag[i]] = 1.0;
  }
  dinv[i] = 1.0 / d;
}

__global__ void pack43_kernel(const double* __restrict__ in4, int64_t nn, double* __restrict__ out3) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  out3[i * 3 + 0] = in4[i * 4 + 0];
  out3[i * 3 + 1] = in4[i * 4 + 1];
  out3[i * 3 + 2] = in4[i * 4 + 2];
}

// ---- K12: reductions ---------------------------------------------------------------------------------
// Every metric kernel leaves NV per-CTA partials; op[k]: 0 sum, 1 max, 2 min.  finalize_kernel combines
// them in CTA order (deterministic).
template <int NV>
__device__ __forceinline__ void block_reduce_ops(double (&v)[NV], const int (&op)[NV], double* __restrict__ partial) {
  __shared__ double s_red[(kT / 32) * NV];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = op[k] == 0 ? warp_sum(v[k]) : (op[k] == 1 ? warp_max(v[k]) : warp_min(v[k]));
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) s_red[wid * NV + k] = v[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double a = s_red[k];
      for (int w = 1; w < kT / 32; ++w) {
        const double b = s_red[w * NV + k];
        a = op[k] == 0 ? a + b : (op[k] == 1 ? fmax(a, b) : fmin(a, b));
      }
      partial[(size_t)blockIdx.x * NV + k] = a;
    }
  }
}

// one warp per output slot: lanes stride over the CTA partials, then a fixed shuffle tree (deterministic)
__global__ void finalize_kernel(const double* __restrict__ partial, int nblocks, int nv, uint32_t opmask2 /*2 bits per slot*/,
                                const int* __restrict__ ops_long, double* __restrict__ out) {
  const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (k >= nv) return;
  const int op = ops_long ? ops_long[k] : (int)((opmask2 >> (2 * k)) & 3u);
  double a = op == 0 ? 0.0 : (op == 1 ? -INFINITY : INFINITY);
  for (int b = lane; b < nblocks; b += 32) {
    const double v = partial[(size_t)b * nv + k];
    a = op == 0 ? a + v : (op == 1 ? fmax(a, v) : fmin(a, v));
  }
  a = op == 0 ? warp_sum(a) : (op == 1 ? warp_max(a) : warp_min(a));
  if (lane == 0) out[k] = a;
}

__device__ __forceinline__ bool in_footprint(double x, double y, const ptfem_footprint& f, double scale) {
  const double dx = x - f.cx, dy = y - f.cy, r = f.r * scale;
  if (f.square) return fabs(dx) < r && fabs(dy) < r;
  return sqrt(dx * dx + dy * dy) < r;
}

struct FpPack {
  ptfem_footprint f[4];
  int n;
};

// nodes with zmin < z (<= zmax unless zmax is NaN): {count, sum, max, min} of the selected field
__global__ void __launch_bounds__(kT) metric_nodes_kernel(const double* __restrict__ xyz, int64_t nn,
                                                          const double* __restrict__ phi, int S, int sys,
                                                          const double* __restrict__ Jn, int field, double zmin,
                                                          double zmax, int mode, FpPack fp, double scale_r,
                                                          double* __restrict__ partial) {
  double v[4] = {0.0, 0.0, -INFINITY, INFINITY};
  const int64_t stride = (int64_t)gridDim.x * kT;
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < nn; i += stride) {
    const double x = xyz[i * 3], y = xyz[i * 3 + 1], z = xyz[i * 3 + 2];
    if (!(z > zmin)) continue;
    if (zmax == zmax && !(z < zmax)) continue;
    if (mode != 0) {
      bool inside = false;
      for (int k = 0; k < fp.n; ++k) inside = inside || in_footprint(x, y, fp.f[k], scale_r);
      if ((mode == 1) != inside) continue;
    }
    double f;
    if (field == 1) {
      f = phi[i * S + sys];
    } else {
      const double jx = Jn[i * 3], jy = Jn[i * 3 + 1], jz = Jn[i * 3 + 2];
      f = field == 0 ? sqrt(jx * jx + jy * jy + jz * jz) : (field == 2 ? fabs(jz) : jz);
    }
    v[0] += 1.0;
    v[1] += f;
    v[2] = fmax(v[2], f);
    v[3] = fmin(v[3], f);
  }
  const int op[4] = {0, 0, 1, 2};
  block_reduce_ops<4>(v, op, partial);
}

// run_layered_sweep.py:704-761 — boundary triangles: J_cell = mean of 3 nodal J, I = sum J_z * area
__global__ void __launch_bounds__(kT) pad_current_kernel(const double* __restrict__ xyz, const int32_t* __restrict__ tris,
                                                         int64_t nb, const double* __restrict__ tri_area,
                                                         const double* __restrict__ Jn, double zmin, ptfem_footprint fp,
                                                         double scale_r, double* __restrict__ partial) {
  double v[3] = {0.0, 0.0, 0.0};
  const int64_t stride = (int64_t)gridDim.x * kT;
  for (int64_t t = (int64_t)blockIdx.x * kT + threadIdx.x; t < nb; t += stride) {
    const int64_t a = tris[t * 3], b = tris[t * 3 + 1], c = tris[t * 3 + 2];
    const double cx = (xyz[a * 3] + xyz[b * 3] + xyz[c * 3]) / 3.0;
    const double cy = (xyz[a * 3 + 1] + xyz[b * 3 + 1] + xyz[c * 3 + 1]) / 3.0;
    const double cz = (xyz[a * 3 + 2] + xyz[b * 3 + 2] + xyz[c * 3 + 2]) / 3.0;
    if (!(cz > zmin) || !in_footprint(cx, cy, fp, scale_r)) continue;
    const double jz = (Jn[a * 3 + 2] + Jn[b * 3 + 2] + Jn[c * 3 + 2]) / 3.0;
    v[0] += jz * tri_area[t];
    v[1] += tri_area[t];
    v[2] += 1.0;
  }
  const int op[3] = {0, 0, 0};
  block_reduce_ops<3>(v, op, partial);
}

// VTK cell->point averaging of the cell-mean potential (vtkCellDataToPointData): unweighted mean over all
// cells (tets, and boundary triangles when include_tris) that use the point.
__global__ void smooth_phi_kernel(const int32_t* __restrict__ n2t_ptr, const int32_t* __restrict__ n2t,
                                  const int32_t* __restrict__ n2b_ptr, const int32_t* __restrict__ n2b,
                                  const int32_t* __restrict__ tets, const int32_t* __restrict__ tris,
                                  const double* __restrict__ phi, int S, int sys, int include_tris, int64_t nn,
                                  const double* __restrict__ xyz, double cx, double cy, double cz, double rad,
                                  double* __restrict__ phis) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  {  // only nodes that a cell inside the ROI can use are needed (rad = largest ROI radius + longest edge)
    const double dx = xyz[3 * i] - cx, dy = xyz[3 * i + 1] - cy, dz = xyz[3 * i + 2] - cz;
    if (dx * dx + dy * dy + dz * dz > rad * rad) return;
  }
  double acc = 0.0;
  int cnt = 0;
  for (int32_t k = n2t_ptr[i]; k < n2t_ptr[i + 1]; ++k) {
    const int64_t t = n2t[k];
    acc += (phi[(int64_t)tets[t * 4] * S + sys] + phi[(int64_t)tets[t * 4 + 1] * S + sys] +
            phi[(int64_t)tets[t * 4 + 2] * S + sys] + phi[(int64_t)tets[t * 4 + 3] * S + sys]) / 4.0;
    ++cnt;
  }
  if (include_tris) {
    for (int32_t k = n2b_ptr[i]; k < n2b_ptr[i + 1]; ++k) {
      const int64_t t = n2b[k];
      acc += (phi[(int64_t)tris[t * 3] * S + sys] + phi[(int64_t)tris[t * 3 + 1] * S + sys] +
              phi[(int64_t)tris[t * 3 + 2] * S + sys]) / 3.0;
      ++cnt;
    }
  }
  phis[i] = cnt > 0 ? acc / (double)cnt : 0.0;
}

// run_layered_sweep.py:765-822,948-959 — cells = tets then tris; per radius multiplier m:
// {n, sum|J_c|, sum|E_c|, n(z>z1), n(z0<z<=z1), n(z<=z0)}
constexpr int kMaxMult = 4;
struct RoiArgs {
  double cen[3];
  double r[kMaxMult];
  int nmult;
  double z0, z1;
};

__global__ void __launch_bounds__(kT) roi_kernel(const double* __restrict__ xyz, const int32_t* __restrict__ tets,
                                                 int64_t nt, const int32_t* __restrict__ tris, int64_t nb,
                                                 const double* __restrict__ phis, const double* __restrict__ Jn,
                                                 RoiArgs a, double* __restrict__ partial) {
  double v[6 * kMaxMult];
#pragma unroll
  for (int k = 0; k < 6 * kMaxMult; ++k) v[k] = 0.0;
  const int64_t stride = (int64_t)gridDim.x * kT;
  const int64_t ncell = nt + nb;
  for (int64_t c = (int64_t)blockIdx.x * kT + threadIdx.x; c < ncell; c += stride) {
    double cx, cy, cz, jm, em;
    if (c < nt) {
      const int32_t* t = tets + c * 4;
      cx = cy = cz = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int64_t n = t[k];
        cx += xyz[n * 3]; cy += xyz[n * 3 + 1]; cz += xyz[n * 3 + 2];
      }
      cx *= 0.25; cy *= 0.25; cz *= 0.25;
      const double dx = cx - a.cen[0], dy = cy - a.cen[1], dz = cz - a.cen[2];
      if (!(sqrt(dx * dx + dy * dy + dz * dz) < a.r[a.nmult - 1])) continue;
      double jx = 0.0, jy = 0.0, jz = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int64_t n = t[k];
        jx += Jn[n * 3]; jy += Jn[n * 3 + 1]; jz += Jn[n * 3 + 2];
      }
      jx *= 0.25; jy *= 0.25; jz *= 0.25;
      jm = sqrt(jx * jx + jy * jy + jz * jz);
      double g[4][3];
      tet_grads(xyz, t, g);
      double ex = 0.0, ey = 0.0, ez = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const double pv = phis[t[k]];
        ex += pv * g[k][0]; ey += pv * g[k][1]; ez += pv * g[k][2];
      }
      em = sqrt(ex * ex + ey * ey + ez * ez);
    } else {
      const int32_t* t = tris + (c - nt) * 3;
      const int64_t n0 = t[0], n1 = t[1], n2 = t[2];
      cx = (xyz[n0 * 3] + xyz[n1 * 3] + xyz[n2 * 3]) / 3.0;
      cy = (xyz[n0 * 3 + 1] + xyz[n1 * 3 + 1] + xyz[n2 * 3 + 1]) / 3.0;
      cz = (xyz[n0 * 3 + 2] + xyz[n1 * 3 + 2] + xyz[n2 * 3 + 2]) / 3.0;
      const double dx = cx - a.cen[0], dy = cy - a.cen[1], dz = cz - a.cen[2];
      if (!(sqrt(dx * dx + dy * dy + dz * dz) < a.r[a.nmult - 1])) continue;
      const double jx = (Jn[n0 * 3] + Jn[n1 * 3] + Jn[n2 * 3]) / 3.0;
      const double jy = (Jn[n0 * 3 + 1] + Jn[n1 * 3 + 1] + Jn[n2 * 3 + 1]) / 3.0;
      const double jz = (Jn[n0 * 3 + 2] + Jn[n1 * 3 + 2] + Jn[n2 * 3 + 2]) / 3.0;
      jm = sqrt(jx * jx + jy * jy + jz * jz);
      // in-plane gradient of the linear interpolant
      const double e1x = xyz[n1 * 3] - xyz[n0 * 3], e1y = xyz[n1 * 3 + 1] - xyz[n0 * 3 + 1], e1z = xyz[n1 * 3 + 2] - xyz[n0 * 3 + 2];
      const double e2x = xyz[n2 * 3] - xyz[n0 * 3], e2y = xyz[n2 * 3 + 1] - xyz[n0 * 3 + 1], e2z = xyz[n2 * 3 + 2] - xyz[n0 * 3 + 2];
      const double nx = e1y * e2z - e1z * e2y, ny = e1z * e2x - e1x * e2z, nz = e1x * e2y - e1y * e2x;
      double n2v = nx * nx + ny * ny + nz * nz;
      n2v = n2v > 0.0 ? n2v : 1.0;
      const double d1 = phis[n1] - phis[n0], d2 = phis[n2] - phis[n0];
      // e2 x n and n x e1
      const double ax = e2y * nz - e2z * ny, ay = e2z * nx - e2x * nz, az = e2x * ny - e2y * nx;
      const double bx = ny * e1z - nz * e1y, by = nz * e1x - nx * e1z, bz = nx * e1y - ny * e1x;
      const double ex = (d1 * ax + d2 * bx) / n2v, ey = (d1 * ay + d2 * by) / n2v, ez = (d1 * az + d2 * bz) / n2v;
      em = sqrt(ex * ex + ey * ey + ez * ez);
    }
    const double dx = cx - a.cen[0], dy = cy - a.cen[1], dz = cz - a.cen[2];
    const double dist = sqrt(dx * dx + dy * dy + dz * dz);
#pragma unroll
    for (int m = 0; m < kMaxMult; ++m) {
      if (m < a.nmult && dist < a.r[m]) {
        v[m * 6 + 0] += 1.0;
        v[m * 6 + 1] += jm;
        v[m * 6 + 2] += em;
        if (cz > a.z1) v[m * 6 + 3] += 1.0;
        else if (cz > a.z0) v[m * 6 + 4] += 1.0;
        else v[m * 6 + 5] += 1.0;
      }
    }
  }
  int op[6 * kMaxMult];
#pragma unroll
  for (int k = 0; k < 6 * kMaxMult; ++k) op[k] = 0;
  block_reduce_ops<6 * kMaxMult>(v, op, partial);
}

// test_step01_baseline.py:76-86 — sums for the least-squares line phi(z) in the centre column
__global__ void __launch_bounds__(kT) column_fit_kernel(const double* __restrict__ xyz, int64_t nn,
                                                        const double* __restrict__ phi, int S, int sys, double cx,
                                                        double cy, double rad, double* __restrict__ partial) {
  double v[6] = {0, 0, 0, 0, 0, 0};
  const int64_t stride = (int64_t)gridDim.x * kT;
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < nn; i += stride) {
    const double dx = xyz[i * 3] - cx, dy = xyz[i * 3 + 1] - cy;
    if (!(hypot(dx, dy) < rad)) continue;
    const double z = xyz[i * 3 + 2], f = phi[i * S + sys];
    v[0] += 1.0; v[1] += z; v[2] += f; v[3] += z * z; v[4] += z * f; v[5] += f * f;
  }
  const int op[6] = {0, 0, 0, 0, 0, 0};
  block_reduce_ops<6>(v, op, partial);
}

__global__ void __launch_bounds__(kT) jstats_kernel(const double* __restrict__ Jn, int64_t nn, double shift,
                                                    double* __restrict__ partial) {
  double v[3] = {0, 0, 0};
  const int64_t stride = (int64_t)gridDim.x * kT;
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < nn; i += stride) {
    const double jx = Jn[i * 3], jy = Jn[i * 3 + 1], jz = Jn[i * 3 + 2];
    const double d = sqrt(jx * jx + jy * jy + jz * jz) - shift;
    v[0] += 1.0; v[1] += d; v[2] += d * d;
  }
  const int op[3] = {0, 0, 0};
  block_reduce_ops<3>(v, op, partial);
}

__global__ void mark_bc_nodes_kernel(const int32_t* __restrict__ tris, const int32_t* __restrict__ bcid, int64_t nb,
                                     int32_t id, uint8_t* __restrict__ flag) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nb || bcid[t] != id) return;
  flag[tris[t * 3]] = 1; flag[tris[t * 3 + 1]] = 1; flag[tris[t * 3 + 2]] = 1;
}

// sum over flagged rows of (K_raw phi - b_neu)_i : the weak-form current through the electrode
__global__ void __launch_bounds__(kT) reaction_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                      const double* __restrict__ val, int VS, int vs,
                                                      const double* __restrict__ phi, int S, int sys,
                                                      const double* __restrict__ b_neu, int BS, int bs,
                                                      const uint8_t* __restrict__ flag, int64_t nn,
                                                      double* __restrict__ partial) {
  double v[1] = {0.0};
  const int64_t stride = (int64_t)gridDim.x * kT;
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < nn; i += stride) {
    if (!flag[i]) continue;
    double acc = 0.0;
    for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) acc += val[(int64_t)k * VS + vs] * phi[(int64_t)col[k] * S + sys];
    v[0] += acc - b_neu[i * BS + bs];
  }
  const int op[1] = {0};
  block_reduce_ops<1>(v, op, partial);
}

// ---- K13 ---------------------------------------------------------------------------------------------
// owner[p] = smallest tet index whose closed cell contains sample point p (atomicMin -> deterministic)
constexpr int kPtsChunk = 512;
__global__ void __launch_bounds__(kT) locate_kernel(const double* __restrict__ xyz, const int32_t* __restrict__ tets,
                                                    int64_t nt, const double* __restrict__ pts, int npts,
                                                    int32_t* __restrict__ owner) {
  __shared__ double sp[kPtsChunk * 3];
  for (int p0 = 0; p0 < npts; p0 += kPtsChunk) {
    const int np = min(kPtsChunk, npts - p0);
    __syncthreads();
    for (int k = threadIdx.x; k < np * 3; k += kT) sp[k] = pts[(int64_t)p0 * 3 + k];
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * kT;
    for (int64_t e = (int64_t)blockIdx.x * kT + threadIdx.x; e < nt; e += stride) {
      double q[4][3];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int64_t n = tets[e * 4 + a];
        q[a][0] = xyz[n * 3]; q[a][1] = xyz[n * 3 + 1]; q[a][2] = xyz[n * 3 + 2];
      }
      double lo[3], hi[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        lo[d] = fmin(fmin(q[0][d], q[1][d]), fmin(q[2][d], q[3][d]));
        hi[d] = fmax(fmax(q[0][d], q[1][d]), fmax(q[2][d], q[3][d]));
      }
      bool have_g = false;
      double g[4][3];
      for (int p = 0; p < np; ++p) {
        const double x = sp[p * 3], y = sp[p * 3 + 1], z = sp[p * 3 + 2];
        if (x < lo[0] || x > hi[0] || y < lo[1] || y > hi[1] || z < lo[2] || z > hi[2]) continue;
        if (!have_g) {
          tet_grads(xyz, tets + e * 4, g);
          have_g = true;
        }
        // barycentric: N_a(x) = N_a(q0) + g_a.(x - q0), N_0(q0) = 1, N_{a>0}(q0) = 0
        const double dx = x - q[0][0], dy = y - q[0][1], dz = z - q[0][2];
        bool inside = true;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const double N = (a == 0 ? 1.0 : 0.0) + g[a][0] * dx + g[a][1] * dy + g[a][2] * dz;
          inside = inside && (N >= -1e-10);
        }
        if (inside) atomicMin(owner + p0 + p, (int32_t)e);
      }
    }
  }
}

__global__ void interp_kernel(const double* __restrict__ xyz, const int32_t* __restrict__ tets,
                              const double* __restrict__ pts, int npts, const int32_t* __restrict__ owner,
                              const double* __restrict__ phi, int S, int sys, int64_t nt, double* __restrict__ out) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npts) return;
  const int64_t e = owner[p];
  if (e < 0 || e >= nt) {
    out[p] = NAN;
    return;
  }
  double g[4][3];
  tet_grads(xyz, tets + e * 4, g);
  const int64_t n0 = tets[e * 4];
  const double dx = pts[p * 3] - xyz[n0 * 3], dy = pts[p * 3 + 1] - xyz[n0 * 3 + 1], dz = pts[p * 3 + 2] - xyz[n0 * 3 + 2];
  double acc = 0.0;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const double N = (a == 0 ? 1.0 : 0.0) + g[a][0] * dx + g[a][1] * dy + g[a][2] * dz;
    acc += N * phi[(int64_t)tets[e * 4 + a] * S + sys];
  }
  out[p] = acc;
}

// second difference along the polyline (non-uniform spacing allowed); end points get 0
__global__ void activating_kernel(const double* __restrict__ pts, const double* __restrict__ v, int npts,
                                  double* __restrict__ af) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npts) return;
  if (p == 0 || p == npts - 1) {
    af[p] = 0.0;
    return;
  }
  const double ax = pts[p * 3] - pts[(p - 1) * 3], ay = pts[p * 3 + 1] - pts[(p - 1) * 3 + 1], az = pts[p * 3 + 2] - pts[(p - 1) * 3 + 2];
  const double bx = pts[(p + 1) * 3] - pts[p * 3], by = pts[(p + 1) * 3 + 1] - pts[p * 3 + 1], bz = pts[(p + 1) * 3 + 2] - pts[p * 3 + 2];
  const double h0 = sqrt(ax * ax + ay * ay + az * az), h1 = sqrt(bx * bx + by * by + bz * bz);
  af[p] = 2.0 * ((v[p + 1] - v[p]) / h1 - (v[p] - v[p - 1]) / h0) / (h0 + h1);
}

int red_grid(ptfem_ctx* ctx, int64_t n) {
  int64_t g = (n + kT - 1) / kT;
  const int64_t cap = (int64_t)ctx->sm_count * 4;
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}

// partial buffer: scratch_d; result -> host
int finish_reduce(ptfem_mesh* m, int grid, int nv, const int* ops, double* out_host) {
  ptfem_ctx* ctx = m->ctx;
  uint32_t mask = 0;
  for (int k = 0; k < nv && k < 16; ++k) mask |= (uint32_t)ops[k] << (2 * k);
  DevBuf<int> d_ops;
  const int* ops_long = nullptr;
  if (nv > 16) {
    PT_TRY(d_ops.alloc(nv));
    PT_CK(cudaMemcpyAsync(d_ops.p, ops, nv * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    ops_long = d_ops.p;
  }
  double* res = m->scratch_d.p + (size_t)grid * nv;
  finalize_kernel<<<1, 32 * nv, 0, ctx->stream>>>(m->scratch_d.p, grid, nv, mask, ops_long, res);
  PT_LAUNCH_CHECK(ctx);
  PT_CK(cudaMemcpyAsync(ctx->h_pinned, res, nv * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  PT_CK(cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < nv; ++k) out_host[k] = ctx->h_pinned[k];
  return PTFEM_OK;
}

int need_partials(ptfem_mesh* m, int grid, int nv) { return m->scratch_d.alloc((size_t)(grid + 1) * nv + 64); }

int check_sys(ptfem_mesh* m, int sys) {
  if (!m->phi.p || m->S < 1) return set_err(PTFEM_ERR_STATE, "no solution on the device (call ptfem_solve first)");
  if (sys < 0 || sys >= m->nsys_user) return set_err(PTFEM_ERR_ARG, "system index %d out of range (0..%d)", sys, m->nsys_user - 1);
  return PTFEM_OK;
}

}  // namespace

// -----------------------------------------------------------------------------------------------------
int ptfem_do_element_fields(ptfem_mesh* m, int sys) {
  ptfem_ctx* ctx = m->ctx;
  PT_TRY(check_sys(m, sys));
  PT_TRY(m->Eelem.alloc((size_t)m->nt * 3));
  PT_TRY(m->Jelem.alloc((size_t)m->nt * 3));
  if (m->nt > 0) {
    const double* sig = m->sigma_tab.p + (size_t)(m->nvalp == 1 ? 0 : sys) * m->nreg;
    element_fields_kernel<<<ceil_div(m->nt, 128), 128, 0, ctx->stream>>>(m->xyz.p, m->tets.p, m->nt, m->phi.p, m->S, sys,
                                                                          m->regidx.p, sig, m->Eelem.p, m->Jelem.p);
    PT_LAUNCH_CHECK(ctx);
  }
  return PTFEM_OK;
}

int ptfem_do_assemble_mass(ptfem_mesh* m);

// Jnode is about to be overwritten: an asynchronous read-back issued by ptfem_recover_current_async must finish first
static int wait_j_copy(ptfem_mesh* m) {
  if (m->j_copy_pending) {
    PT_CK(cudaStreamWaitEvent(m->ctx->stream, m->ctx->ev_j_copied, 0));
    m->j_copy_pending = false;
  }
  return PTFEM_OK;
}

int ptfem_do_recover(ptfem_mesh* m, int sys, int method) {
  ptfem_ctx* ctx = m->ctx;
  if (m->J_all_valid && m->J_all_method == method) {   // already there (ptfem_recover_current_batch)
    PT_TRY(check_sys(m, sys));
    m->J_sys = sys;
    m->J_from_all = true;
    return PTFEM_OK;
  }
  m->J_all_valid = false;
  m->J_from_all = false;
  PT_TRY(ptfem_do_element_fields(m, sys));
  PT_TRY(m->Jnode.alloc((size_t)m->nn * 3));
  const int grid = ceil_div(m->nn, 128);
  if (method == PTFEM_RECOVER_LUMPED || method == PTFEM_RECOVER_AVERAGE) {
    PT_TRY(wait_j_copy(m));
    recover_gather_kernel<<<grid, 128, 0, ctx->stream>>>(m->n2t_ptr.p, m->n2t.p, m->vol.p, m->Jelem.p, m->mlump.p, m->nn,
                                                         method == PTFEM_RECOVER_LUMPED ? 1 : 2, m->Jnode.p);
    PT_LAUNCH_CHECK(ctx);
    m->J_sys = sys;
    return PTFEM_OK;
  }
  if (method != PTFEM_RECOVER_L2) return set_err(PTFEM_ERR_ARG, "unknown recovery method %d", method);
  // consistent-mass L2 projection: three right-hand sides on the mesh pattern, Jacobi-PCG
  const bool fresh = !(m->mval.p && m->mval.n >= (size_t)m->nnz);
  PT_TRY(ptfem_do_assemble_mass(m));
  PT_TRY(m->mdinv.alloc(m->nn));
  if (fresh) {
    mass_dinv_kernel<<<grid, 128, 0, ctx->stream>>>(m->diag.p, m->mval.p, m->nn, m->mdinv.p);
    PT_LAUNCH_CHECK(ctx);
  }
  PT_TRY(m->mrhs.alloc((size_t)m->nn * 4));
  PT_TRY(m->mx.alloc((size_t)m->nn * 4));
  recover_gather_kernel<<<grid, 128, 0, ctx->stream>>>(m->n2t_ptr.p, m->n2t.p, m->vol.p, m->Jelem.p, m->mlump.p, m->nn, 0,
                                                       m->mrhs.p);
  PT_LAUNCH_CHECK(ctx);
  PT_CK(cudaMemsetAsync(m->mx.p, 0, (size_t)m->nn * 4 * sizeof(double), ctx->stream));
  LinSys A;
  A.nn = m->nn; A.nnz = m->nnz; A.rowptr = m->rowptr.p; A.col = m->col.p; A.val = m->mval.p; A.VS = 1; A.S = 4;
  A.dinv = m->mdinv.p; A.b = m->mrhs.p; A.stream_rows = 0;
  if (!m->has_rowperm) {   // the tile geometry depends on the pattern only: the mass solve can stream too
    A.stream_rows = m->stream_rows;
    A.stream_cap = m->stream_cap;
  }
  ptfem_solve_opts o;
  ptfem_solve_opts_default(&o);
  o.precond = PTFEM_PRECOND_JACOBI;
  o.rtol = 1e-11;   // mass matrix, kappa(D^-1 M) <= 5: ~25 iterations; leaves J at ~1e-11, far inside the 1e-4 bar
  o.maxit = 2000;
  o.check_every = 10;
  o.use_graph = 0;
  ptfem_solve_stats st;
  PT_TRY(pcg_solve(ctx, A, m->work3, o, m->mx.p, &st));
  m->mass_iters = st.iterations;
  PT_TRY(wait_j_copy(m));
  pack43_kernel<<<grid, 128, 0, ctx->stream>>>(m->mx.p, m->nn, m->Jnode.p);
  PT_LAUNCH_CHECK(ctx);
  m->J_sys = sys;
  return PTFEM_OK;
}

// nodal current of system sys on the device: a block of the batched recovery, or the single-system buffer
static const double* j_of(ptfem_mesh* m, int sys) {
  if (m->J_all_valid && m->Jall.p && sys >= 0 && sys < m->nsys_user) return m->Jall.p + (size_t)sys * m->nn * 3;
  if (m->Jnode.p && m->J_sys == sys && !m->J_from_all) return m->Jnode.p;
  return nullptr;
}
const double* ptfem_current_ptr(ptfem_mesh* m) { return m->J_sys >= 0 ? j_of(m, m->J_sys) : nullptr; }
static int need_J(ptfem_mesh* m, int sys) {
  if (!j_of(m, sys))
    return set_err(PTFEM_ERR_STATE, "nodal current of system %d has not been recovered (ptfem_recover_current)", sys);
  return PTFEM_OK;
}

int ptfem_do_metric_nodes(ptfem_mesh* m, int sys, int field, double zmin, double zmax, int mode, const ptfem_footprint* fp,
                          int nfp, double scale_r, double out[4]) {
  ptfem_ctx* ctx = m->ctx;
  PT_TRY(check_sys(m, sys));
  if (field != 1) PT_TRY(need_J(m, sys));
  if (nfp < 0 || nfp > 4) return set_err(PTFEM_ERR_ARG, "at most 4 footprints");
  FpPack pk;
  pk.n = nfp;
  for (int k = 0; k < nfp; ++k) pk.f[k] = fp[k];
  const int grid = red_grid(ctx, m->nn);
  PT_TRY(need_partials(m, grid, 4));
  metric_nodes_kernel<<<grid, kT, 0, ctx->stream>>>(m->xyz.p, m->nn, m->phi.p, m->S, sys, j_of(m, sys), field, zmin, zmax, mode,
                                                    pk, scale_r, m->scratch_d.p);
  PT_LAUNCH_CHECK(ctx);
  const int ops[4] = {0, 0, 1, 2};
  return finish_reduce(m, grid, 4, ops, out);
}

int ptfem_do_metric_pad_current(ptfem_mesh* m, int sys, double zmin, const ptfem_footprint* fp, double scale_r, double out[3]) {
  ptfem_ctx* ctx = m->ctx;
  PT_TRY(check_sys(m, sys));
  PT_TRY(need_J(m, sys));
  const int grid = red_grid(ctx, m->nb);
  PT_TRY(need_partials(m, grid, 3));
  pad_current_kernel<<<grid, kT, 0, ctx->stream>>>(m->xyz.p, m->tris.p, m->nb, m->tri_area.p, j_of(m, sys), zmin, *fp, scale_r,
                                                   m->scratch_d.p);
  PT_LAUNCH_CHECK(ctx);
  const int ops[3] = {0, 0, 0};
  return finish_reduce(m, grid, 3, ops, out);
}

int ptfem_do_metric_roi(ptfem_mesh* m, int sys, const double cen[3], double r0, const double* mult, int nmult, double z0,
                        double z1, int include_tris, double* out) {
  ptfem_ctx* ctx = m->ctx;
  PT_TRY(check_sys(m, sys));
  PT_TRY(need_J(m, sys));
  if (nmult < 1 || nmult > kMaxMult) return set_err(PTFEM_ERR_ARG, "1..4 radius multipliers");
  for (int k = 1; k < nmult; ++k)
    if (!(mult[k] >= mult[k - 1])) return set_err(PTFEM_ERR_ARG, "radius multipliers must be non-decreasing");
  PT_TRY(m->phis.alloc(m->nn));
  smooth_phi_kernel<<<ceil_div(m->nn, 128), 128, 0, ctx->stream>>>(m->n2t_ptr.p, m->n2t.p, m->n2b_ptr.p, m->n2b.p, m->tets.p,
                                                                    m->tris.p, m->phi.p, m->S, sys, include_tris, m->nn,
                                                                    m->xyz.p, cen[0], cen[1], cen[2],
                                                                    r0 * mult[nmult - 1] + 1.01 * m->h_max, m->phis.p);
  PT_LAUNCH_CHECK(ctx);
  RoiArgs a;
  for (int k = 0; k < 3; ++k) a.cen[k] = cen[k];
  for (int k = 0; k < kMaxMult; ++k) a.r[k] = r0 * mult[k < nmult ? k : nmult - 1];
  a.nmult = nmult;
  a.z0 = z0;
  a.z1 = z1;
  const int64_t ncell = m->nt + (include_tris ? m->nb : 0);
  const int grid = red_grid(ctx, ncell);
  constexpr int NV = 6 * kMaxMult;
  PT_TRY(need_partials(m, grid, NV));
  roi_kernel<<<grid, kT, 0, ctx->stream>>>(m->xyz.p, m->tets.p, m->nt, m->tris.p, include_tris ? m->nb : 0, m->phis.p,
                                           j_of(m, sys), a, m->scratch_d.p);
  PT_LAUNCH_CHECK(ctx);
  int ops[NV];
  for (int k = 0; k < NV; ++k) ops[k] = 0;
  double res[NV];
  PT_TRY(finish_reduce(m, grid, NV, ops, res));
  for (int k = 0; k < nmult * 6; ++k) out[k] = res[k];
  return PTFEM_OK;
}

int ptfem_do_metric_column_fit(ptfem_mesh* m, int sys, double cx, double cy, double rad, double out[6]) {
  ptfem_ctx* ctx = m->ctx;
  PT_TRY(check_sys(m, sys));
  const int grid = red_grid(ctx, m->nn);
  PT_TRY(need_partials(m, grid, 6));
  column_fit_kernel<<<grid, kT, 0, ctx->stream>>>(m->xyz.p, m->nn, m->phi.p, m->S, sys, cx, cy, rad, m->scratch_d.p);
  PT_LAUNCH_CHECK(ctx);
  const int ops[6] = {0, 0, 0, 0, 0, 0};
  return finish_reduce(m, grid, 6, ops, out);
}

int ptfem_do_metric_jstats(ptfem_mesh* m, int sys, double shift, double out[3]) {
  ptfem_ctx* ctx = m->ctx;
  PT_TRY(need_J(m, sys));
  const int grid = red_grid(ctx, m->nn);
  PT_TRY(need_partials(m, grid, 3));
  jstats_kernel<<<grid, kT, 0, ctx->stream>>>(j_of(m, sys), m->nn, shift, m->scratch_d.p);
  PT_LAUNCH_CHECK(ctx);
  const int ops[3] = {0, 0, 0};
  return finish_reduce(m, grid, 3, ops, out);
}

int ptfem_do_metric_reaction(ptfem_mesh* m, int sys, int32_t bcid, double* current) {
  ptfem_ctx* ctx = m->ctx;
  PT_TRY(check_sys(m, sys));
  if (!m->b_neu.p) return set_err(PTFEM_ERR_STATE, "boundary conditions have not been set");
  DevBuf<uint8_t> flag;
  PT_TRY(flag.alloc(m->nn));
  PT_CK(cudaMemsetAsync(flag.p, 0, m->nn, ctx->stream));
  if (m->nb > 0) {
    mark_bc_nodes_kernel<<<ceil_div(m->nb, 256), 256, 0, ctx->stream>>>(m->tris.p, m->bcid.p, m->nb, bcid, flag.p);
    PT_LAUNCH_CHECK(ctx);
  }
  const int grid = red_grid(ctx, m->nn);
  PT_TRY(need_partials(m, grid, 1));
  const int VS = m->nvalp, vs = VS == 1 ? 0 : sys;
  const int BS = m->nrhsp, bs = BS == 1 ? 0 : sys;
  reaction_kernel<<<grid, kT, 0, ctx->stream>>>(m->rowptr.p, m->col.p, m->val_raw.p, VS, vs, m->phi.p, m->S, sys, m->b_neu.p, BS,
                                                bs, flag.p, m->nn, m->scratch_d.p);
  PT_LAUNCH_CHECK(ctx);
  const int ops[1] = {0};
  int rc = finish_reduce(m, grid, 1, ops, current);
  return rc;  // finish_reduce synchronised the stream: flag may go out of scope
}

int ptfem_do_sample_polyline(ptfem_mesh* m, int sys, int64_t npts, const double* pts, double* phi_out, double* af_out) {
  ptfem_ctx* ctx = m->ctx;
  PT_TRY(check_sys(m, sys));
  if (npts <= 0) return PTFEM_OK;
  if (npts > (1 << 24)) return set_err(PTFEM_ERR_ARG, "too many sample points");
  DevBuf<double> d_pts, d_v, d_af;
  DevBuf<int32_t> owner;
  PT_TRY(d_pts.alloc(npts * 3));
  PT_TRY(d_v.alloc(npts));
  PT_TRY(d_af.alloc(npts));
  PT_TRY(owner.alloc(npts));
  PT_CK(cudaMemcpyAsync(d_pts.p, pts, npts * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  PT_TRY(fill_i32(ctx, owner.p, 0x7fffffff, npts));
  if (m->nt > 0) {
    int grid = ceil_div(m->nt, kT);
    if (grid > ctx->sm_count * 8) grid = ctx->sm_count * 8;
    locate_kernel<<<grid, kT, 0, ctx->stream>>>(m->xyz.p, m->tets.p, m->nt, d_pts.p, (int)npts, owner.p);
    PT_LAUNCH_CHECK(ctx);
  }
  interp_kernel<<<ceil_div(npts, 128), 128, 0, ctx->stream>>>(m->xyz.p, m->tets.p, d_pts.p, (int)npts, owner.p, m->phi.p, m->S,
                                                               sys, m->nt, d_v.p);
  PT_LAUNCH_CHECK(ctx);
  activating_kernel<<<ceil_div(npts, 128), 128, 0, ctx->stream>>>(d_pts.p, d_v.p, (int)npts, d_af.p);
  PT_LAUNCH_CHECK(ctx);
  if (phi_out) PT_CK(cudaMemcpyAsync(phi_out, d_v.p, npts * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (af_out) PT_CK(cudaMemcpyAsync(af_out, d_af.p, npts * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  PT_CK(cudaStreamSynchronize(ctx->stream));
  return PTFEM_OK;
}

// =====================================================================================================
// Sweep forms: nodal currents of every system in two launches, and any number of metric reductions in one
// pass per kind with a single read-back.  The reference runs one ElmerSolver + pyvista extraction per sweep
// point (step02_electrodes/run_sweep.py:301-341, step03_ankle_layers/run_layered_sweep.py:1061-1062); here the
// S configurations solved together are also post-processed together (vectors are [nn][S], system fastest).
// =====================================================================================================
namespace {

// ---- K10/K11 batched: J_e of all S systems -> Je[nt][3][S]; LPT = S/2 lanes per tet (one 16-byte gather per node) ----
template <int S>
__global__ void __launch_bounds__(256) element_J_all_kernel(const double* __restrict__ xyz, const int32_t* __restrict__ tets,
                                                            int64_t nt, const double* __restrict__ phi,
                                                            const uint8_t* __restrict__ regidx, const double* __restrict__ sigma,
                                                            int VS, int nreg, double* __restrict__ Je) {
  constexpr int LPT = S >= 2 ? S / 2 : 1;
  constexpr int NV = S >= 2 ? 2 : 1;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t e = gid / LPT;
  const int lane = (int)(gid % LPT);
  if (e >= nt) return;
  double g[4][3];
  tet_grads(xyz, tets + e * 4, g);
  double ex[NV], ey[NV], ez[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) ex[v] = ey[v] = ez[v] = 0.0;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const double* pp = phi + (int64_t)tets[e * 4 + a] * S + NV * lane;
    double pv[NV];
    if constexpr (NV == 2) {
      const double2 t = *reinterpret_cast<const double2*>(pp);
      pv[0] = t.x;
      pv[1] = t.y;
    } else {
      pv[0] = pp[0];
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      ex[v] -= pv[v] * g[a][0];
      ey[v] -= pv[v] * g[a][1];
      ez[v] -= pv[v] * g[a][2];
    }
  }
  const int r = regidx[e];
  double sg[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) sg[v] = sigma[(size_t)(VS == 1 ? 0 : NV * lane + v) * nreg + r];
  double* out = Je + (size_t)e * 3 * S + NV * lane;
  if constexpr (NV == 2) {
    *reinterpret_cast<double2*>(out) = make_double2(sg[0] * ex[0], sg[1] * ex[1]);
    *reinterpret_cast<double2*>(out + S) = make_double2(sg[0] * ey[0], sg[1] * ey[1]);
    *reinterpret_cast<double2*>(out + 2 * S) = make_double2(sg[0] * ez[0], sg[1] * ez[1]);
  } else {
    out[0] = sg[0] * ex[0];
    out[S] = sg[0] * ey[0];
    out[2 * S] = sg[0] * ez[0];
  }
}

// node gather (sorted node -> tet lists: fixed order): Jall[sys][i][k] ; mode 1 lumped, 2 unweighted average
template <int S>
__global__ void __launch_bounds__(256) recover_gather_all_kernel(const int32_t* __restrict__ n2t_ptr, const int32_t* __restrict__ n2t,
                                                                 const double* __restrict__ vol, const double* __restrict__ Je,
                                                                 const double* __restrict__ mlump, int64_t nn, int mode, int nsys,
                                                                 double* __restrict__ Jall) {
  constexpr int LPN = S >= 2 ? S / 2 : 1;
  constexpr int NV = S >= 2 ? 2 : 1;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = gid / LPN;
  const int lane = (int)(gid % LPN);
  if (i >= nn) return;
  const int32_t b = n2t_ptr[i], e = n2t_ptr[i + 1];
  double acc[3][NV];
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[k][v] = 0.0;
  for (int32_t q = b; q < e; ++q) {
    const int64_t t = n2t[q];
    const double w = mode == 2 ? 1.0 : 0.25 * vol[t];
    const double* src = Je + (size_t)t * 3 * S + NV * lane;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if constexpr (NV == 2) {
        const double2 jv = *reinterpret_cast<const double2*>(src + k * S);
        acc[k][0] += w * jv.x;
        acc[k][1] += w * jv.y;
      } else {
        acc[k][0] += w * src[k * S];
      }
    }
  }
  double d = mode == 1 ? mlump[i] : (double)(e - b);
  d = d > 0.0 ? 1.0 / d : 0.0;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int sys = NV * lane + v;
    if (sys >= nsys) continue;
    double* out = Jall + ((size_t)sys * nn + i) * 3;
    out[0] = acc[0][v] * d;
    out[1] = acc[1][v] * d;
    out[2] = acc[2][v] * d;
  }
}

template <int S>
int recover_batch_t(ptfem_mesh* m, int mode) {
  ptfem_ctx* ctx = m->ctx;
  constexpr int L = S >= 2 ? S / 2 : 1;
  if (m->nt > 0) {
    element_J_all_kernel<S><<<ceil_div(m->nt * L, 256), 256, 0, ctx->stream>>>(m->xyz.p, m->tets.p, m->nt, m->phi.p, m->regidx.p,
                                                                               m->sigma_tab.p, m->nvalp, m->nreg, m->Jelem_all.p);
    PT_LAUNCH_CHECK(ctx);
  }
  recover_gather_all_kernel<S><<<ceil_div(m->nn * L, 256), 256, 0, ctx->stream>>>(m->n2t_ptr.p, m->n2t.p, m->vol.p, m->Jelem_all.p,
                                                                                  m->mlump.p, m->nn, mode, m->nsys_user, m->Jall.p);
  PT_LAUNCH_CHECK(ctx);
  return PTFEM_OK;
}

// ---- K12 batched -----------------------------------------------------------------------------------------
// single-precision tet centroids: prefilter of the ROI scans (a cell within the margin is re-tested in double)
__global__ void tet_centroid_kernel(const double* __restrict__ xyz, const int32_t* __restrict__ tets, int64_t nt,
                                    float4* __restrict__ cen) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nt) return;
  double c[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int64_t n = tets[e * 4 + a];
    c[0] += xyz[n * 3];
    c[1] += xyz[n * 3 + 1];
    c[2] += xyz[n * 3 + 2];
  }
  cen[e] = make_float4((float)(0.25 * c[0]), (float)(0.25 * c[1]), (float)(0.25 * c[2]), 0.f);
}

// bounding sphere of every run of kT consecutive tet centroids (one block per run): an ROI scan tests the run first and skips it
// whole - on meshes whose tets are stored in spatial order (structured, extruded, anything a mesher emits region by region) a 5 mm
// ROI touches a few hundred of the 77 k runs of the 20 M-tet slab, so a request reads ~1 MB instead of the 316 MB centroid array
__global__ void __launch_bounds__(kT) tet_chunk_kernel(const float4* __restrict__ cen, int64_t nt, float4* __restrict__ chunk) {
  __shared__ float s_a[kT], s_b[kT], s_c[kT];
  const int tid = threadIdx.x;
  const int64_t e = (int64_t)blockIdx.x * kT + tid;
  const int n = (int)min((int64_t)kT, nt - (int64_t)blockIdx.x * kT);
  const float4 q = e < nt ? cen[e] : make_float4(0.f, 0.f, 0.f, 0.f);
  s_a[tid] = q.x; s_b[tid] = q.y; s_c[tid] = q.z;
  __syncthreads();
  for (int o = kT / 2; o > 0; o >>= 1) {
    if (tid < o) { s_a[tid] += s_a[tid + o]; s_b[tid] += s_b[tid + o]; s_c[tid] += s_c[tid + o]; }
    __syncthreads();
  }
  const float cx = s_a[0] / n, cy = s_b[0] / n, cz = s_c[0] / n;
  __syncthreads();
  const float dx = q.x - cx, dy = q.y - cy, dz = q.z - cz;
  s_a[tid] = e < nt ? sqrtf(dx * dx + dy * dy + dz * dz) : 0.f;
  __syncthreads();
  for (int o = kT / 2; o > 0; o >>= 1) {
    if (tid < o) s_a[tid] = fmaxf(s_a[tid], s_a[tid + o]);
    __syncthreads();
  }
  if (tid == 0) chunk[blockIdx.x] = make_float4(cx, cy, cz, s_a[0] * 1.0001f + 1e-30f);
}

struct BatchReq {   // device copy of one request (+ where its inputs / partials live)
  ptfem_metric_req r;
  int32_t slot;     // index among the requests of its kind
  int32_t pad_;
};

// blockIdx.y = request (of kind NODES); partial[(slot*gridDim.x + blockIdx.x)*4 ..]
__global__ void __launch_bounds__(kT) metric_nodes_batch_kernel(const double* __restrict__ xyz, int64_t nn,
                                                                const double* __restrict__ phi, int S,
                                                                const double* __restrict__ Jall, const BatchReq* __restrict__ reqs,
                                                                const int32_t* __restrict__ idx, double* __restrict__ partial) {
  const BatchReq& q = reqs[idx[blockIdx.y]];
  const ptfem_metric_req& r = q.r;
  const double* Jn = Jall + (size_t)r.sys * nn * 3;
  double v[4] = {0.0, 0.0, -INFINITY, INFINITY};
  const int64_t stride = (int64_t)gridDim.x * kT;
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < nn; i += stride) {
    const double z = xyz[i * 3 + 2];
    if (!(z > r.zmin)) continue;
    if (r.zmax == r.zmax && !(z < r.zmax)) continue;
    if (r.mode != 0) {
      const double x = xyz[i * 3], y = xyz[i * 3 + 1];
      bool inside = false;
      for (int k = 0; k < r.nfp; ++k) inside = inside || in_footprint(x, y, r.fp[k], r.scale_r);
      if ((r.mode == 1) != inside) continue;
    }
    double f;
    if (r.field == 1) {
      f = phi[i * S + r.sys];
    } else {
      const double jx = Jn[i * 3], jy = Jn[i * 3 + 1], jz = Jn[i * 3 + 2];
      f = r.field == 0 ? sqrt(jx * jx + jy * jy + jz * jz) : (r.field == 2 ? fabs(jz) : jz);
    }
    v[0] += 1.0;
    v[1] += f;
    v[2] = fmax(v[2], f);
    v[3] = fmin(v[3], f);
  }
  const int op[4] = {0, 0, 1, 2};
  block_reduce_ops<4>(v, op, partial + (size_t)blockIdx.y * gridDim.x * 4);
}

__global__ void __launch_bounds__(kT) pad_current_batch_kernel(const double* __restrict__ xyz, const int32_t* __restrict__ tris,
                                                               int64_t nb, const double* __restrict__ tri_area, int64_t nn,
                                                               const double* __restrict__ Jall, const BatchReq* __restrict__ reqs,
                                                               const int32_t* __restrict__ idx, double* __restrict__ partial) {
  const ptfem_metric_req& r = reqs[idx[blockIdx.y]].r;
  const double* Jn = Jall + (size_t)r.sys * nn * 3;
  double v[3] = {0.0, 0.0, 0.0};
  const int64_t stride = (int64_t)gridDim.x * kT;
  for (int64_t t = (int64_t)blockIdx.x * kT + threadIdx.x; t < nb; t += stride) {
    const int64_t a = tris[t * 3], b = tris[t * 3 + 1], c = tris[t * 3 + 2];
    const double cx = (xyz[a * 3] + xyz[b * 3] + xyz[c * 3]) / 3.0;
    const double cy = (xyz[a * 3 + 1] + xyz[b * 3 + 1] + xyz[c * 3 + 1]) / 3.0;
    const double cz = (xyz[a * 3 + 2] + xyz[b * 3 + 2] + xyz[c * 3 + 2]) / 3.0;
    if (!(cz > r.zmin) || !in_footprint(cx, cy, r.fp[0], r.scale_r)) continue;
    const double jz = (Jn[a * 3 + 2] + Jn[b * 3 + 2] + Jn[c * 3 + 2]) / 3.0;
    v[0] += jz * tri_area[t];
    v[1] += tri_area[t];
    v[2] += 1.0;
  }
  const int op[3] = {0, 0, 0};
  block_reduce_ops<3>(v, op, partial + (size_t)blockIdx.y * gridDim.x * 3);
}

// VTK point smoothing for every ROI request (blockIdx.y): phis[slot][i] for the nodes an in-ROI cell can use
// bounding sphere (double precision) of every run of 128 consecutive nodes - one block of the smoothing kernel - so that blocks far
// from the ROI leave after one 32-byte read instead of reading their nodes' coordinates
__global__ void __launch_bounds__(128) node_chunk_kernel(const double* __restrict__ xyz, int64_t nn, double* __restrict__ chunk) {
  __shared__ double s_a[128], s_b[128], s_c[128];
  const int tid = threadIdx.x;
  const int64_t i = (int64_t)blockIdx.x * 128 + tid;
  const int n = (int)min((int64_t)128, nn - (int64_t)blockIdx.x * 128);
  const double x = i < nn ? xyz[3 * i] : 0.0, y = i < nn ? xyz[3 * i + 1] : 0.0, z = i < nn ? xyz[3 * i + 2] : 0.0;
  s_a[tid] = x; s_b[tid] = y; s_c[tid] = z;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (tid < o) { s_a[tid] += s_a[tid + o]; s_b[tid] += s_b[tid + o]; s_c[tid] += s_c[tid + o]; }
    __syncthreads();
  }
  const double cx = s_a[0] / n, cy = s_b[0] / n, cz = s_c[0] / n;
  __syncthreads();
  s_a[tid] = i < nn ? sqrt((x - cx) * (x - cx) + (y - cy) * (y - cy) + (z - cz) * (z - cz)) : 0.0;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (tid < o) s_a[tid] = fmax(s_a[tid], s_a[tid + o]);
    __syncthreads();
  }
  if (tid == 0) {
    chunk[4 * (size_t)blockIdx.x] = cx; chunk[4 * (size_t)blockIdx.x + 1] = cy; chunk[4 * (size_t)blockIdx.x + 2] = cz;
    chunk[4 * (size_t)blockIdx.x + 3] = s_a[0] * (1.0 + 1e-12);
  }
}

__global__ void smooth_phi_batch_kernel(const int32_t* __restrict__ n2t_ptr, const int32_t* __restrict__ n2t,
                                        const int32_t* __restrict__ n2b_ptr, const int32_t* __restrict__ n2b,
                                        const int32_t* __restrict__ tets, const int32_t* __restrict__ tris,
                                        const double* __restrict__ phi, int S, int64_t nn, const double* __restrict__ xyz,
                                        double h_max, const BatchReq* __restrict__ reqs, const int32_t* __restrict__ idx,
                                        const double* __restrict__ nchunk, double* __restrict__ phis_all) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  const ptfem_metric_req& r = reqs[idx[blockIdx.y]].r;
  const double rad = r.r0 * r.mult[r.nmult - 1] + 1.01 * h_max;
  {   // the block's nodes (blockDim.x == 128 consecutive ones) all lie outside: nothing to do
    const double ex = nchunk[4 * (size_t)blockIdx.x] - r.cen[0], ey = nchunk[4 * (size_t)blockIdx.x + 1] - r.cen[1],
                 ez = nchunk[4 * (size_t)blockIdx.x + 2] - r.cen[2], lim = (rad + nchunk[4 * (size_t)blockIdx.x + 3]) * (1.0 + 1e-12);
    if (ex * ex + ey * ey + ez * ez > lim * lim) return;
  }
  const double dx = xyz[3 * i] - r.cen[0], dy = xyz[3 * i + 1] - r.cen[1], dz = xyz[3 * i + 2] - r.cen[2];
  if (dx * dx + dy * dy + dz * dz > rad * rad) return;
  const int sys = r.sys;
  double acc = 0.0;
  int cnt = 0;
  for (int32_t k = n2t_ptr[i]; k < n2t_ptr[i + 1]; ++k) {
    const int64_t t = n2t[k];
    acc += (phi[(int64_t)tets[t * 4] * S + sys] + phi[(int64_t)tets[t * 4 + 1] * S + sys] +
            phi[(int64_t)tets[t * 4 + 2] * S + sys] + phi[(int64_t)tets[t * 4 + 3] * S + sys]) / 4.0;
    ++cnt;
  }
  if (r.include_tris) {
    for (int32_t k = n2b_ptr[i]; k < n2b_ptr[i + 1]; ++k) {
      const int64_t t = n2b[k];
      acc += (phi[(int64_t)tris[t * 3] * S + sys] + phi[(int64_t)tris[t * 3 + 1] * S + sys] +
              phi[(int64_t)tris[t * 3 + 2] * S + sys]) / 3.0;
      ++cnt;
    }
  }
  phis_all[(size_t)blockIdx.y * nn + i] = cnt > 0 ? acc / (double)cnt : 0.0;
}

// the per-cell part of roi_kernel for one cell known to be a candidate; adds to v[6*kMaxMult]
__device__ __forceinline__ void roi_cell_accumulate(const double* __restrict__ xyz, const int32_t* __restrict__ tets, int64_t nt,
                                                    const int32_t* __restrict__ tris, int64_t c, const double* __restrict__ phis,
                                                    const double* __restrict__ Jn, const ptfem_metric_req& a, double (&v)[6 * kMaxMult]) {
  double cx, cy, cz, jm, em;
  const double rmax = a.r0 * a.mult[a.nmult - 1];
  if (c < nt) {
    const int32_t* t = tets + c * 4;
    cx = cy = cz = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t n = t[k];
      cx += xyz[n * 3]; cy += xyz[n * 3 + 1]; cz += xyz[n * 3 + 2];
    }
    cx *= 0.25; cy *= 0.25; cz *= 0.25;
    const double dx = cx - a.cen[0], dy = cy - a.cen[1], dz = cz - a.cen[2];
    if (!(sqrt(dx * dx + dy * dy + dz * dz) < rmax)) return;
    double jx = 0.0, jy = 0.0, jz = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t n = t[k];
      jx += Jn[n * 3]; jy += Jn[n * 3 + 1]; jz += Jn[n * 3 + 2];
    }
    jx *= 0.25; jy *= 0.25; jz *= 0.25;
    jm = sqrt(jx * jx + jy * jy + jz * jz);
    double g[4][3];
    tet_grads(xyz, t, g);
    double ex = 0.0, ey = 0.0, ez = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double pv = phis[t[k]];
      ex += pv * g[k][0]; ey += pv * g[k][1]; ez += pv * g[k][2];
    }
    em = sqrt(ex * ex + ey * ey + ez * ez);
  } else {
    const int32_t* t = tris + (c - nt) * 3;
    const int64_t n0 = t[0], n1 = t[1], n2 = t[2];
    cx = (xyz[n0 * 3] + xyz[n1 * 3] + xyz[n2 * 3]) / 3.0;
    cy = (xyz[n0 * 3 + 1] + xyz[n1 * 3 + 1] + xyz[n2 * 3 + 1]) / 3.0;
    cz = (xyz[n0 * 3 + 2] + xyz[n1 * 3 + 2] + xyz[n2 * 3 + 2]) / 3.0;
    const double dx = cx - a.cen[0], dy = cy - a.cen[1], dz = cz - a.cen[2];
    if (!(sqrt(dx * dx + dy * dy + dz * dz) < rmax)) return;
    const double jx = (Jn[n0 * 3] + Jn[n1 * 3] + Jn[n2 * 3]) / 3.0;
    const double jy = (Jn[n0 * 3 + 1] + Jn[n1 * 3 + 1] + Jn[n2 * 3 + 1]) / 3.0;
    const double jz = (Jn[n0 * 3 + 2] + Jn[n1 * 3 + 2] + Jn[n2 * 3 + 2]) / 3.0;
    jm = sqrt(jx * jx + jy * jy + jz * jz);
    const double e1x = xyz[n1 * 3] - xyz[n0 * 3], e1y = xyz[n1 * 3 + 1] - xyz[n0 * 3 + 1], e1z = xyz[n1 * 3 + 2] - xyz[n0 * 3 + 2];
    const double e2x = xyz[n2 * 3] - xyz[n0 * 3], e2y = xyz[n2 * 3 + 1] - xyz[n0 * 3 + 1], e2z = xyz[n2 * 3 + 2] - xyz[n0 * 3 + 2];
    const double nx = e1y * e2z - e1z * e2y, ny = e1z * e2x - e1x * e2z, nz = e1x * e2y - e1y * e2x;
    double n2v = nx * nx + ny * ny + nz * nz;
    n2v = n2v > 0.0 ? n2v : 1.0;
    const double d1 = phis[n1] - phis[n0], d2 = phis[n2] - phis[n0];
    const double ax = e2y * nz - e2z * ny, ay = e2z * nx - e2x * nz, az = e2x * ny - e2y * nx;
    const double bx = ny * e1z - nz * e1y, by = nz * e1x - nx * e1z, bz = nx * e1y - ny * e1x;
    const double ex = (d1 * ax + d2 * bx) / n2v, ey = (d1 * ay + d2 * by) / n2v, ez = (d1 * az + d2 * bz) / n2v;
    em = sqrt(ex * ex + ey * ey + ez * ez);
  }
  const double dx = cx - a.cen[0], dy = cy - a.cen[1], dz = cz - a.cen[2];
  const double dist = sqrt(dx * dx + dy * dy + dz * dz);
#pragma unroll
  for (int mi = 0; mi < kMaxMult; ++mi) {
    if (mi < a.nmult && dist < a.r0 * a.mult[mi]) {
      v[mi * 6 + 0] += 1.0;
      v[mi * 6 + 1] += jm;
      v[mi * 6 + 2] += em;
      if (cz > a.z1) v[mi * 6 + 3] += 1.0;
      else if (cz > a.z0) v[mi * 6 + 4] += 1.0;
      else v[mi * 6 + 5] += 1.0;
    }
  }
}

// blockIdx.y = ROI request.  Tets are prefiltered by their single-precision centroid (16 bytes per tet instead of the
// 4 node ids + 4 gathered coordinates), boundary triangles are tested directly.
__global__ void __launch_bounds__(kT) roi_batch_kernel(const double* __restrict__ xyz, const int32_t* __restrict__ tets, int64_t nt,
                                                       const int32_t* __restrict__ tris, int64_t nb, const float4* __restrict__ tcen,
                                                       const float4* __restrict__ tchunk, int64_t nn, const double* __restrict__ phis_all, const double* __restrict__ Jall,
                                                       const BatchReq* __restrict__ reqs, const int32_t* __restrict__ idx,
                                                       double* __restrict__ partial) {
  const ptfem_metric_req& a = reqs[idx[blockIdx.y]].r;
  const double* phis = phis_all + (size_t)blockIdx.y * nn;
  const double* Jn = Jall + (size_t)a.sys * nn * 3;
  double v[6 * kMaxMult];
#pragma unroll
  for (int k = 0; k < 6 * kMaxMult; ++k) v[k] = 0.0;
  const double rmax = a.r0 * a.mult[a.nmult - 1];
  // margin of the float prefilter: a few ulps of the coordinates' magnitude
  const float cxf = (float)a.cen[0], cyf = (float)a.cen[1], czf = (float)a.cen[2];
  const float big = fmaxf(fmaxf(fabsf(cxf), fabsf(cyf)), fabsf(czf)) + (float)rmax;
  const float rf = (float)rmax * 1.0001f + 64.f * 1.1920929e-7f * big;
  const float rf2 = rf * rf;
  const int64_t stride = (int64_t)gridDim.x * kT;
  const int64_t nchunk = (nt + kT - 1) / kT;
  for (int64_t ch = blockIdx.x; ch < nchunk; ch += gridDim.x) {      // thread t of the block owns tet ch * kT + t
    const float4 sp = __ldg(tchunk + ch);
    const float ex = sp.x - cxf, ey = sp.y - cyf, ez = sp.z - czf;
    const float lim = (rf + sp.w) * 1.00001f;
    if (ex * ex + ey * ey + ez * ez > lim * lim) continue;            // the whole run lies outside (same answer in every thread)
    const int64_t c = ch * kT + threadIdx.x;
    if (c >= nt) continue;
    const float4 q = __ldg(tcen + c);
    const float dx = q.x - cxf, dy = q.y - cyf, dz = q.z - czf;
    if (dx * dx + dy * dy + dz * dz > rf2) continue;
    roi_cell_accumulate(xyz, tets, nt, tris, c, phis, Jn, a, v);
  }
  if (a.include_tris)
    for (int64_t c = nt + (int64_t)blockIdx.x * kT + threadIdx.x; c < nt + nb; c += stride)
      roi_cell_accumulate(xyz, tets, nt, tris, c, phis, Jn, a, v);
  int op[6 * kMaxMult];
#pragma unroll
  for (int k = 0; k < 6 * kMaxMult; ++k) op[k] = 0;
  block_reduce_ops<6 * kMaxMult>(v, op, partial + (size_t)blockIdx.y * gridDim.x * 6 * kMaxMult);
}

// one warp per (request, slot): fixed lane-strided order over the CTA partials, fixed shuffle tree
__global__ void finalize_batch_kernel(const double* __restrict__ partial, int nblocks, int nv, int nreq_kind, uint32_t opmask2,
                                      const int32_t* __restrict__ idx, double* __restrict__ out) {
  const int w = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (w >= nreq_kind * nv) return;
  const int rq = w / nv, k = w % nv;
  const int op = (int)((opmask2 >> (2 * (k < 16 ? k : 15))) & 3u);
  const double* p = partial + (size_t)rq * nblocks * nv;
  double a = op == 0 ? 0.0 : (op == 1 ? -INFINITY : INFINITY);
  for (int b = lane; b < nblocks; b += 32) {
    const double x = p[(size_t)b * nv + k];
    a = op == 0 ? a + x : (op == 1 ? fmax(a, x) : fmin(a, x));
  }
  a = op == 0 ? warp_sum(a) : (op == 1 ? warp_max(a) : warp_min(a));
  if (lane == 0) out[(size_t)idx[rq] * PTFEM_METRIC_OUT_STRIDE + k] = a;
}

}  // namespace

int ptfem_do_recover_batch(ptfem_mesh* m, int method) {
  PT_TRY(check_sys(m, 0));
  if (method == PTFEM_RECOVER_L2) {
    // consistent-mass projection needs a PCG solve per system: system by system into the blocks of Jall
    PT_TRY(m->Jall.alloc((size_t)m->nsys_user * m->nn * 3));
    PT_TRY(wait_j_copy(m));
    for (int s = 0; s < m->nsys_user; ++s) {
      m->J_all_valid = false;
      PT_TRY(ptfem_do_recover(m, s, PTFEM_RECOVER_L2));
      PT_CK(cudaMemcpyAsync(m->Jall.p + (size_t)s * m->nn * 3, m->Jnode.p, (size_t)m->nn * 3 * sizeof(double), cudaMemcpyDeviceToDevice,
                            m->ctx->stream));
    }
  } else if (method == PTFEM_RECOVER_LUMPED || method == PTFEM_RECOVER_AVERAGE) {
    if (m->nvalp != 1 && m->nvalp != m->S) return set_err(PTFEM_ERR_STATE, "value sets and systems out of step");
    PT_TRY(m->Jall.alloc((size_t)m->nsys_user * m->nn * 3));
    PT_TRY(m->Jelem_all.alloc((size_t)(m->nt > 0 ? m->nt : 1) * 3 * m->S));
    PT_TRY(wait_j_copy(m));
    const int mode = method == PTFEM_RECOVER_LUMPED ? 1 : 2;
    switch (m->S) {
      case 1: PT_TRY(recover_batch_t<1>(m, mode)); break;
      case 2: PT_TRY(recover_batch_t<2>(m, mode)); break;
      case 4: PT_TRY(recover_batch_t<4>(m, mode)); break;
      case 8: PT_TRY(recover_batch_t<8>(m, mode)); break;
      case 16: PT_TRY(recover_batch_t<16>(m, mode)); break;
      default: return set_err(PTFEM_ERR_STATE, "unsupported system count %d", m->S);
    }
  } else {
    return set_err(PTFEM_ERR_ARG, "unknown recovery method %d", method);
  }
  m->J_all_valid = true;
  m->J_all_method = method;
  m->J_sys = 0;
  m->J_from_all = true;
  return PTFEM_OK;
}

int ptfem_do_metrics_batch(ptfem_mesh* m, int32_t nreq, const ptfem_metric_req* req, double* out) {
  ptfem_ctx* ctx = m->ctx;
  PT_TRY(check_sys(m, 0));
  std::vector<BatchReq> h(nreq);
  std::vector<int32_t> idx[3];
  bool need_j = false;
  for (int q = 0; q < nreq; ++q) {
    const ptfem_metric_req& r = req[q];
    if (r.kind < 0 || r.kind > 2) return set_err(PTFEM_ERR_ARG, "request %d: unknown kind %d", q, r.kind);
    PT_TRY(check_sys(m, r.sys));
    if (r.kind == PTFEM_METRIC_NODES) {
      if (r.field < 0 || r.field > 3 || r.mode < 0 || r.mode > 2) return set_err(PTFEM_ERR_ARG, "request %d: bad field / mode", q);
      if (r.mode != 0 && (r.nfp < 1 || r.nfp > 2)) return set_err(PTFEM_ERR_ARG, "request %d: 1..2 footprints for mode 1/2", q);
      need_j = need_j || r.field != 1;
    } else if (r.kind == PTFEM_METRIC_ROI) {
      if (r.nmult < 1 || r.nmult > kMaxMult) return set_err(PTFEM_ERR_ARG, "request %d: 1..4 radius multipliers", q);
      for (int k = 1; k < r.nmult; ++k)
        if (!(r.mult[k] >= r.mult[k - 1])) return set_err(PTFEM_ERR_ARG, "request %d: radius multipliers must be non-decreasing", q);
      need_j = true;
    } else {
      need_j = true;
    }
    h[q].r = r;
    h[q].slot = (int32_t)idx[r.kind].size();
    h[q].pad_ = 0;
    idx[r.kind].push_back(q);
  }
  if (need_j && !(m->J_all_valid && m->Jall.p))
    return set_err(PTFEM_ERR_STATE, "ptfem_metrics_batch needs the nodal currents of every system (ptfem_recover_current_batch)");
  const double* Jall = m->Jall.p;
  if (!m->has_tcen && !idx[2].empty()) {
    PT_TRY(m->tcen.alloc((size_t)(m->nt > 0 ? m->nt : 1) * 4));
    if (m->nt > 0) {
      tet_centroid_kernel<<<ceil_div(m->nt, 256), 256, 0, ctx->stream>>>(m->xyz.p, m->tets.p, m->nt, reinterpret_cast<float4*>(m->tcen.p));
      PT_LAUNCH_CHECK(ctx);
    }
    PT_TRY(m->tchunk.alloc((size_t)(ceil_div(m->nt, kT) > 0 ? ceil_div(m->nt, kT) : 1) * 4));
    if (m->nt > 0) {
      tet_chunk_kernel<<<ceil_div(m->nt, kT), kT, 0, ctx->stream>>>(reinterpret_cast<const float4*>(m->tcen.p), m->nt,
                                                                    reinterpret_cast<float4*>(m->tchunk.p));
      PT_LAUNCH_CHECK(ctx);
    }
    PT_TRY(m->nchunk.alloc((size_t)ceil_div(m->nn, 128) * 4));
    node_chunk_kernel<<<ceil_div(m->nn, 128), 128, 0, ctx->stream>>>(m->xyz.p, m->nn, m->nchunk.p);
    PT_LAUNCH_CHECK(ctx);
    m->has_tcen = true;
  }
  // device copies: requests, per-kind index lists, partials, results
  DevBuf<BatchReq> d_req;
  DevBuf<int32_t> d_idx;
  PT_TRY(d_req.alloc(nreq));
  PT_TRY(d_idx.alloc(nreq));
  std::vector<int32_t> flat;
  int off[3];
  for (int k = 0; k < 3; ++k) {
    off[k] = (int)flat.size();
    flat.insert(flat.end(), idx[k].begin(), idx[k].end());
  }
  PT_CK(cudaMemcpyAsync(d_req.p, h.data(), sizeof(BatchReq) * nreq, cudaMemcpyHostToDevice, ctx->stream));
  PT_CK(cudaMemcpyAsync(d_idx.p, flat.data(), sizeof(int32_t) * nreq, cudaMemcpyHostToDevice, ctx->stream));
  const int gn = red_grid(ctx, m->nn), gb = red_grid(ctx, m->nb), gc = red_grid(ctx, m->nt + m->nb);
  constexpr int NVR = 6 * kMaxMult;
  const size_t np0 = idx[0].size() * (size_t)gn * 4, np1 = idx[1].size() * (size_t)gb * 3, np2 = idx[2].size() * (size_t)gc * NVR;
  const size_t nres = (size_t)nreq * PTFEM_METRIC_OUT_STRIDE;
  PT_TRY(m->scratch_d.alloc(np0 + np1 + np2 + nres + 64));
  double* part0 = m->scratch_d.p;
  double* part1 = part0 + np0;
  double* part2 = part1 + np1;
  double* res = part2 + np2;
  PT_CK(cudaMemsetAsync(res, 0, nres * sizeof(double), ctx->stream));
  if (!idx[0].empty()) {
    metric_nodes_batch_kernel<<<dim3(gn, (unsigned)idx[0].size()), kT, 0, ctx->stream>>>(m->xyz.p, m->nn, m->phi.p, m->S, Jall, d_req.p,
                                                                                       d_idx.p + off[0], part0);
    PT_LAUNCH_CHECK(ctx);
    finalize_batch_kernel<<<ceil_div((int64_t)idx[0].size() * 4 * 32, 256), 256, 0, ctx->stream>>>(part0, gn, 4, (int)idx[0].size(),
                                                                                                    0u | (0u << 2) | (1u << 4) | (2u << 6),
                                                                                                    d_idx.p + off[0], res);
    PT_LAUNCH_CHECK(ctx);
  }
  if (!idx[1].empty()) {
    pad_current_batch_kernel<<<dim3(gb, (unsigned)idx[1].size()), kT, 0, ctx->stream>>>(m->xyz.p, m->tris.p, m->nb, m->tri_area.p, m->nn,
                                                                                      Jall, d_req.p, d_idx.p + off[1], part1);
    PT_LAUNCH_CHECK(ctx);
    finalize_batch_kernel<<<ceil_div((int64_t)idx[1].size() * 3 * 32, 256), 256, 0, ctx->stream>>>(part1, gb, 3, (int)idx[1].size(), 0u,
                                                                                                    d_idx.p + off[1], res);
    PT_LAUNCH_CHECK(ctx);
  }
  if (!idx[2].empty()) {
    PT_TRY(m->phis_all.alloc(idx[2].size() * (size_t)m->nn));
    smooth_phi_batch_kernel<<<dim3(ceil_div(m->nn, 128), (unsigned)idx[2].size()), 128, 0, ctx->stream>>>(
        m->n2t_ptr.p, m->n2t.p, m->n2b_ptr.p, m->n2b.p, m->tets.p, m->tris.p, m->phi.p, m->S, m->nn, m->xyz.p, m->h_max, d_req.p,
        d_idx.p + off[2], m->nchunk.p, m->phis_all.p);
    PT_LAUNCH_CHECK(ctx);
    roi_batch_kernel<<<dim3(gc, (unsigned)idx[2].size()), kT, 0, ctx->stream>>>(m->xyz.p, m->tets.p, m->nt, m->tris.p, m->nb,
                                                                                 reinterpret_cast<const float4*>(m->tcen.p),
                                                                                 reinterpret_cast<const float4*>(m->tchunk.p), m->nn,
                                                                                 m->phis_all.p, Jall, d_req.p, d_idx.p + off[2], part2);
    PT_LAUNCH_CHECK(ctx);
    finalize_batch_kernel<<<ceil_div((int64_t)idx[2].size() * NVR * 32, 256), 256, 0, ctx->stream>>>(part2, gc, NVR, (int)idx[2].size(), 0u,
                                                                                                      d_idx.p + off[2], res);
    PT_LAUNCH_CHECK(ctx);
  }
  PT_CK(cudaMemcpyAsync(out, res, nres * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  PT_CK(cudaStreamSynchronize(ctx->stream));   // (the request / index buffers go back to the allocator here)
  return PTFEM_OK;
}
