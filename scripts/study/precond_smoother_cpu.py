"""CPU study: Jacobi term of the coarse-grid preconditioner replaced by a short polynomial smoother (extra SpMVs per iteration)."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from pelvistim_fem_b200 import meshgen
from oracle import fem_oracle as fo, coarse_oracle as cz
size = sys.argv[1] if len(sys.argv) > 1 else "M"
m = meshgen.synth_slab(size, contact_enabled=False)
SIG = {1: 0.35, 2: 0.04, 3: 0.001, 4: 0.005, 5: 0.005}
K_raw = fo.assemble_stiffness(m.nodes, m.tets, m.region, SIG)
is_dir, val = fo.dirichlet_nodes(m.tris, m.bcid, [(102, 0.0)], m.nn)
K, b = fo.apply_dirichlet_symmetric(K_raw, fo.neumann_rhs(m.nodes, m.tris, m.bcid, [(101, 15.975)]), is_dir, val)
K = K.tocsr()
M = cz.CoarsePreconditioner(K, m.nodes, is_dir, coarse_nodes=300, extra_levels=-1)
print("nn", m.nn, "levels", M.nlev)
t = time.time(); x, it = cz.pcg(K, b, M.apply, rtol=1e-10); print("jacobi + grids: %d iterations (%.1fs)" % (it, time.time() - t), flush=True)
dinv = M.dinv
def coarse(r):
    z = np.zeros_like(r)
    for Z, B in zip(M.Z, M.B):
        rc = Z.T @ r
        z += Z @ (B @ rc if B.ndim == 2 else B * rc)
    return z
# damped-Jacobi smoothing steps around the same coarse correction, symmetric: z = S r + C r with S = two-step damped Jacobi
for omega in (0.6, 0.8, 1.0):
    def apply2(r, w=omega):
        z1 = w * dinv * r
        z2 = z1 + w * dinv * (r - K @ z1)          # two sweeps = degree-1 polynomial in D^-1 A (symmetric)
        return z2 + coarse(r)
    x, it = cz.pcg(K, b, apply2, rtol=1e-10); print("2-sweep damped Jacobi (omega %.1f) + grids: %d iterations, 1 extra product each" % (omega, it), flush=True)
# Chebyshev degree 2 on [lmax/ratio, lmax] of D^-1 A
lmax = 2.0
for ratio in (4.0, 8.0):
    a, bnd = lmax / ratio, lmax
    theta, delta = 0.5 * (bnd + a), 0.5 * (bnd - a)
    def cheb(r, deg=2):
        # standard Chebyshev iteration for D^-1 A z = D^-1 r from z = 0
        sigma = theta / delta
        rho = 1.0 / sigma
        res = dinv * r
        d = res / theta
        z = d.copy()
        for _ in range(deg - 1):
            res = dinv * (r - K @ z)
            rho_new = 1.0 / (2.0 * sigma - rho)
            d = rho_new * rho * d + 2.0 * rho_new / delta * res
            rho = rho_new
            z = z + d
        return z + coarse(r)
    x, it = cz.pcg(K, b, cheb, rtol=1e-10); print("Chebyshev degree 2 (ratio %.0f) + grids: %d iterations, 1 extra product each" % (ratio, it), flush=True)
