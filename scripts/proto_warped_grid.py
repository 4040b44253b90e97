"""CPU study: coarse grids whose z-planes are snapped onto the material interfaces (the slab's conductivity jumps 350x
between skin and muscle) vs the uniform grids of csrc/coarse.cu.  The warp z -> z' is piecewise linear and monotone, the
grid is uniform in z' - so only the table builder (coarse_table_kernel) would change on the device.
Usage: python scripts/proto_warped_grid.py M|L"""
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

# material interfaces from the mesh itself: z-extent of every region
cent_z = mesh.nodes[mesh.tets].mean(axis=1)[:, 2]
faces = sorted({round(float(mesh.nodes[mesh.tets[mesh.region == r]][:, :, 2].max()), 9) for r in np.unique(mesh.region)} |
               {round(float(mesh.nodes[:, 2].min()), 9)})
print(size, "nn", nn, "interfaces (mm):", [round(1e3 * f, 3) for f in faces], flush=True)

def run(name, nodes_for_grid, level_weight=None):
    t = time.time()
    M = cz.CoarsePreconditioner(K, nodes_for_grid, isd, coarse_nodes=2000, extra_levels=-1, level_weight=level_weight)
    x, it = cz.pcg(K, b, M.apply, rtol=1e-10, maxit=3000)
    print("%-60s levels %d coarse %d its %4d  (%.0fs)" % (name, M.nlev, M.coarse_unknowns, it, time.time() - t), flush=True)

run("uniform grids (coarse.cu)", mesh.nodes)
lo, hi = mesh.nodes.min(axis=0), mesh.nodes.max(axis=0)
base = cz.choose_grid(lo, hi, 2000.0)
nz = int(base[2])
# warp: the interfaces are mapped onto coarsest-grid planes; thin layers get at least one coarsest cell each
def warp(zplanes_idx):
    zp = np.asarray(faces); tp = lo[2] + (hi[2] - lo[2]) * np.asarray(zplanes_idx, dtype=float) / nz
    X = mesh.nodes.copy(); X[:, 2] = np.interp(mesh.nodes[:, 2], zp, tp); return X
nf = len(faces)
for idx in ([0, nz - 3, nz - 2, nz - 1, nz][-nf:] if nf <= 5 else None,
            [0, nz - 4, nz - 2, nz - 1, nz][-nf:] if nf <= 5 else None):
    if idx is None or len(idx) != nf:
        continue
    idx[0] = 0
    run("z-planes %s of %d on the interfaces" % (idx, nz), warp(idx))
