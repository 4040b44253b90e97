"""CPU study (scipy + oracle/coarse_oracle.py) of variants of the coarse-grid preconditioner on the bench workload: PCG
iteration counts to rtol 1e-10 for level weights and for a polynomial (Chebyshev in D^-1 K) fine-level smoother in place
of plain Jacobi.  Usage: python scripts/proto_precond_variants.py M [coarse_nodes]"""
import sys, time, importlib
import numpy as np, scipy.sparse as sp
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import fem_oracle as fo
import coarse_oracle as cz
mg = importlib.import_module("pelvistim-fem_b200.meshgen")
import bench

size = sys.argv[1] if len(sys.argv) > 1 else "M"
cn = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
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
if isinstance(isd, tuple):
    isd, val = isd
else:
    val = np.zeros(nn)
K, b = fo.apply_dirichlet_symmetric(K, b, isd, val)
K = K.tocsr()
t0 = time.time()
M = cz.CoarsePreconditioner(K, mesh.nodes, isd, coarse_nodes=cn, extra_levels=-1)
print(size, "nn", nn, "levels", M.nlev, "coarse", M.coarse_unknowns, "setup %.1fs" % (time.time() - t0), flush=True)
dinv = M.dinv

def coarse_part(r, w_exact=1.0, w_bpx=1.0):
    z = np.zeros_like(r)
    for Z, B in zip(M.Z, M.B):
        rc = Z.T @ r
        z += (w_exact * (Z @ (B @ rc))) if B.ndim == 2 else (w_bpx * (Z @ (B * rc)))
    return z

# largest eigenvalue of D^-1 K (power iteration)
v = np.random.default_rng(0).standard_normal(nn)
for _ in range(30):
    v = dinv * (K @ v); lam = np.linalg.norm(v); v /= lam
print("lambda_max(D^-1 K) ~ %.3f" % lam, flush=True)

def cheb_smoother(r, deg, ratio):
    """z = p(D^-1 K) D^-1 r, p = Chebyshev polynomial preconditioner of degree deg-1 for the interval [lmax/ratio, lmax]."""
    lmax, lmin = 1.05 * lam, 1.05 * lam / ratio
    theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
    sigma = theta / delta
    rho = 1.0 / sigma
    z = np.zeros_like(r)
    res = r.copy()
    d = dinv * res / theta
    for k in range(deg):
        z += d
        if k == deg - 1:
            break
        res = res - K @ d
        rho_new = 1.0 / (2.0 * sigma - rho)
        d = rho_new * rho * d + 2.0 * rho_new / delta * (dinv * res)
        rho = rho_new
    return z

def run(name, apply, extra_spmv=0):
    t = time.time()
    x, it = cz.pcg(K, b, apply, rtol=1e-10, maxit=5000)
    # cost model from the B200 launch list on L, 8 RHS: 0.88 ms per iteration, +0.33 per extra SpMM and vector pass
    print("%-44s its %4d   model ms on L (x%.2f per it): %6.1f   (%.0fs)" % (name, it, 0.88 + 0.33 * extra_spmv, it * (0.88 + 0.33 * extra_spmv), time.time() - t), flush=True)
    return it

def coarse_levels_w(r, w):
    z = np.zeros_like(r)
    for Z, B, wl in zip(M.Z, M.B, w):
        rc = Z.T @ r
        z += wl * (Z @ (B @ rc if B.ndim == 2 else B * rc))
    return z

if len(sys.argv) > 3 and sys.argv[3] == "levels":      # per-level weights, finest first, exact level last
    run("baseline", lambda r: dinv * r + coarse_part(r))
    for w in ((0.7, 0.7, 0.7), (0.5, 0.5, 0.5), (0.5, 0.7, 1.0), (0.35, 0.5, 0.7), (0.5, 0.5, 1.0), (1.0, 0.5, 0.5), (0.35, 0.35, 0.35)):
        w = w[-M.nlev:]
        run("level weights " + "/".join("%.2f" % v for v in w), lambda r, w=w: dinv * r + coarse_levels_w(r, w))
    sys.exit(0)
run("baseline  D^-1 + exact + BPX", lambda r: dinv * r + coarse_part(r))
for we, wb in ((1.0, 0.5), (1.0, 2.0), (0.5, 1.0), (2.0, 1.0), (1.5, 1.5), (0.7, 0.7)):
    run("weights exact %.1f bpx %.1f" % (we, wb), lambda r, we=we, wb=wb: dinv * r + coarse_part(r, we, wb))
for deg, ratio in ((2, 4.0), (2, 8.0), (3, 8.0), (3, 16.0)):
    run("cheb deg %d ratio %g + coarse" % (deg, ratio), lambda r, deg=deg, ratio=ratio: cheb_smoother(r, deg, ratio) + coarse_part(r), extra_spmv=deg - 1)
