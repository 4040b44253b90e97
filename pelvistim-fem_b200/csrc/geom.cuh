// geom.cuh — linear-tetrahedron geometry shared by assembly.cu and post.cu.
#pragma once
#include <stdint.h>

// symmetric index of local pair (a,b), a,b in 0..3 -> 0..9
static __device__ __forceinline__ int sym10(int a, int b) {
  // rows: (0,0)=0 (0,1)=1 (0,2)=2 (0,3)=3 (1,1)=4 (1,2)=5 (1,3)=6 (2,2)=7 (2,3)=8 (3,3)=9
  const int lo = a < b ? a : b, hi = a < b ? b : a;
  return lo * 4 - (lo * (lo - 1)) / 2 + (hi - lo);
}

// shape-function gradients of a linear tet; returns signed volume
static __device__ __forceinline__ double tet_grads(const double* __restrict__ xyz, const int32_t* __restrict__ tet,
                                            double g[4][3]) {
  double p[4][3];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int64_t n = tet[a];
    p[a][0] = xyz[n * 3 + 0];
    p[a][1] = xyz[n * 3 + 1];
    p[a][2] = xyz[n * 3 + 2];
  }
  const double a0 = p[1][0] - p[0][0], a1 = p[1][1] - p[0][1], a2 = p[1][2] - p[0][2];
  const double b0 = p[2][0] - p[0][0], b1 = p[2][1] - p[0][1], b2 = p[2][2] - p[0][2];
  const double c0 = p[3][0] - p[0][0], c1 = p[3][1] - p[0][1], c2 = p[3][2] - p[0][2];
  // cofactors: grad N1 = (b x c)/det, grad N2 = (c x a)/det, grad N3 = (a x b)/det
  const double bc0 = b1 * c2 - b2 * c1, bc1 = b2 * c0 - b0 * c2, bc2 = b0 * c1 - b1 * c0;
  const double ca0 = c1 * a2 - c2 * a1, ca1 = c2 * a0 - c0 * a2, ca2 = c0 * a1 - c1 * a0;
  const double ab0 = a1 * b2 - a2 * b1, ab1 = a2 * b0 - a0 * b2, ab2 = a0 * b1 - a1 * b0;
  const double det = a0 * bc0 + a1 * bc1 + a2 * bc2;
  const double inv = det != 0.0 ? 1.0 / det : 0.0;
  g[1][0] = bc0 * inv; g[1][1] = bc1 * inv; g[1][2] = bc2 * inv;
  g[2][0] = ca0 * inv; g[2][1] = ca1 * inv; g[2][2] = ca2 * inv;
  g[3][0] = ab0 * inv; g[3][1] = ab1 * inv; g[3][2] = ab2 * inv;
#pragma unroll
  for (int k = 0; k < 3; ++k) g[0][k] = -(g[1][k] + g[2][k] + g[3][k]);
  return det / 6.0;
}

