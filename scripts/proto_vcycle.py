"""CPU study for the next preconditioner: a symmetric multigrid V-cycle on the SAME nested regular grids as csrc/coarse.cu
(Galerkin operators A_l = Z_l^T K Z_l - 27-point-like stencils on regular grids - instead of only their diagonals;
damped-Jacobi smoothing on the mesh and on every grid; exact coarsest solve) against the additive preconditioner.
Cost model per PCG iteration on the B200 (L, 8 RHS): additive 0.88 ms; V(1,1) = + 2 SpMM passes over the mesh matrix
(residual after pre-smoothing, post-smoothing) ~ + 0.66 ms + grid-level work.
Usage: python scripts/proto_vcycle.py M|L"""
import sys, time, importlib
import numpy as np, scipy.sparse as sp
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import fem_oracle as fo
import coarse_oracle as cz
mg = importlib.import_module("pelvistim-fem_b200.meshgen")
import bench

size = sys.argv[1] if len(sys.argv) > 1 else "M"
mesh = mg.synth_slab(size, contact_enabled=False)
confs = bench.sweep_definition(mesh, 8)
K = fo.assemble_stiffness(mesh.nodes, mesh.tets, mesh.region, bench.SIGMA).tocsr()
nn = mesh.nn
b = np.zeros(nn)
c = confs[3]
tr = mesh.tris[c["tris"]]; p = mesh.nodes[tr]
ar = 0.5 * np.linalg.norm(np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]), axis=1)
for a in range(3):
    np.add.at(b, tr[:, a], bench.I_INJECT / c["area"] * ar / 3)
isd = fo.dirichlet_nodes(mesh.tris, mesh.bcid, [(102, 0.0)], nn)
isd, val = isd if isinstance(isd, tuple) else (isd, np.zeros(nn))
K, b = fo.apply_dirichlet_symmetric(K, b, isd, val)
K = K.tocsr()
t0 = time.time()
M = cz.CoarsePreconditioner(K, mesh.nodes, isd, coarse_nodes=2000, extra_levels=-1)
print(size, "nn", nn, "levels", M.nlev, "grid unknowns", [Z.shape[1] for Z in M.Z], "setup %.0fs" % (time.time() - t0), flush=True)
x, it = cz.pcg(K, b, M.apply, rtol=1e-10)
print("additive (current, weighted)                     its %4d   model %.1f ms" % (it, it * 0.88), flush=True)

# Galerkin operators on the grids, level 0 = finest grid; empty grid nodes get a unit diagonal
A = []
for Z in M.Z:
    E = (Z.T @ K @ Z).tocsr()
    d = E.diagonal()
    E = E + sp.diags((d <= 0.0).astype(float))
    A.append(E.tocsr())
print("grid operators: nnz per row", [round(E.nnz / E.shape[0], 1) for E in A], flush=True)
# grid-to-grid prolongations P[l]: level l (coarser) -> level l-1, from Z_l = Z_{l-1} P_l solved in the least-squares sense is
# overkill: build them geometrically like the tests do
lo, hi = mesh.nodes.min(axis=0), mesh.nodes.max(axis=0)
base = cz.choose_grid(lo, hi, 2000.0)
P = [None]
for l in range(1, M.nlev):
    nf = base * (1 << (M.nlev - l))
    ext = np.where(hi - lo > 0.0, hi - lo, 1.0) * (1.0 + 1e-12)
    ax = [lo[d] + ext[d] * np.arange(nf[d] + 1) / nf[d] for d in range(3)]
    Zg, Yg, Xg = np.meshgrid(ax[2], ax[1], ax[0], indexing="ij")
    pts = np.stack([Xg.ravel(), Yg.ravel(), Zg.ravel()], axis=1)
    P.append(cz.interpolation(pts, np.ones(pts.shape[0]), lo, hi, nf // 2))
Binv = np.linalg.inv(A[-1].toarray())

def lam_max(Aop, dinv, n):
    v = np.random.default_rng(1).standard_normal(n)
    lam = 1.0
    for _ in range(25):
        v = dinv * (Aop @ v); lam = np.linalg.norm(v); v /= lam
    return lam
dK = 1.0 / K.diagonal()
dA = [1.0 / E.diagonal() for E in A]
lK = lam_max(K, dK, nn)
lA = [lam_max(E, d, E.shape[0]) for E, d in zip(A, dA)]
print("lambda_max(D^-1 A): mesh %.2f grids %s" % (lK, [round(v, 2) for v in lA]), flush=True)

def vcycle_grid(l, r, nu, om):
    if l == M.nlev - 1:
        return Binv @ r
    E, d = A[l], dA[l]
    w = om / lA[l]
    x = w * d * r
    for _ in range(nu - 1):
        x += w * d * (r - E @ x)
    x += P[l + 1] @ vcycle_grid(l + 1, P[l + 1].T @ (r - E @ x), nu, om)
    for _ in range(nu):
        x += w * d * (r - E @ x)
    return x

def vcycle(r, nu, om):
    w = om / lK
    x = w * dK * r
    for _ in range(nu - 1):
        x += w * dK * (r - K @ x)
    x += M.Z[0] @ vcycle_grid(0, M.Z[0].T @ (r - K @ x), nu, om)
    for _ in range(nu):
        x += w * dK * (r - K @ x)
    return x

for nu, om in ((1, 1.0), (1, 4.0 / 3.0), (1, 1.6), (2, 4.0 / 3.0)):
    t = time.time()
    x, it = cz.pcg(K, b, lambda r: vcycle(r, nu, om), rtol=1e-10, maxit=500)
    extra = 2 * nu        # SpMM passes over the mesh matrix beyond the CG's own: nu-1 + 1 residual + nu post = 2 nu
    print("V(%d,%d) damped Jacobi omega %.2f/lambda_max            its %4d   model %.1f ms (%.2f ms per it)   (%.0fs)" %
          (nu, nu, om, it, it * (0.88 + 0.33 * extra + 0.05), 0.88 + 0.33 * extra + 0.05, time.time() - t), flush=True)
