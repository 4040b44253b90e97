"""CPU study: block PCG (O'Leary) against 8 independent PCGs for the bench sweep (8 Neumann patches, one matrix),
with the Jacobi + coarse-grid preconditioner of oracle/coarse_oracle.py.   python scripts/proto_block_cg.py [size]"""
import sys, time
sys.path.insert(0, ".")
import numpy as np, scipy.sparse as sp
import bench
import pelvistim_fem_b200  # noqa
from pelvistim_fem_b200 import meshgen
from oracle import c_oracle as co, coarse_oracle as cor

size = sys.argv[1] if len(sys.argv) > 1 else "M"
co.use_all_cores()
mesh = meshgen.synth_slab(size, contact_enabled=False)
confs = bench.sweep_definition(mesh, 8, 0)
cpu = bench.CpuSweep(mesh, confs, "coarse")
cpu.shared(0)
cs = cpu.cs
K = sp.csr_matrix((cs.val, cs.col, cs.rowptr), shape=(cs.nn, cs.nn))
B = np.stack([cpu.rhs(c) for c in confs], axis=1)
P = cor.CoarsePreconditioner(K, mesh.nodes, cs.isdir.astype(bool))
print("levels", P.nlev, "coarse", P.coarse_unknowns)
minv = lambda R: np.stack([P.apply(R[:, j]) for j in range(R.shape[1])], axis=1)
rtol = 1e-10
bn = np.linalg.norm(B, axis=0)
# independent PCGs
its = []
for j in range(8):
    x, it = cor.pcg(K, B[:, j], P.apply, rtol=rtol)
    its.append(it)
print("independent PCG iterations:", its, "max", max(its))
# block PCG
X = np.zeros_like(B); R = B.copy(); Z = minv(R); Pd = Z.copy(); rho = Z.T @ R
it = 0
while it < 500:
    Q = K @ Pd
    alpha = np.linalg.solve(Pd.T @ Q, rho)
    X += Pd @ alpha
    R -= Q @ alpha
    it += 1
    rel = np.linalg.norm(R, axis=0) / bn
    if rel.max() <= rtol:
        break
    Z = minv(R)
    rho_new = Z.T @ R
    beta = np.linalg.solve(rho, rho_new)
    Pd = Z + Pd @ beta
    rho = rho_new
    if it % 5 == 0:
        print(it, "max rel", rel.max(), "cond(PtAP) %.2e" % np.linalg.cond(Pd.T @ (K @ Pd)))
true = np.linalg.norm(B - K @ X, axis=0) / bn
print("block PCG iterations:", it, "recurrence max rel", rel.max(), "true max rel", true.max())
