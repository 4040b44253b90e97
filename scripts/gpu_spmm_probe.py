"""Short 8-RHS PCG run on the L mesh (for ncu captures of the multi-RHS kernels)."""
import sys
sys.path.insert(0, ".")
import numpy as np
import pelvistim_fem_b200
from pelvistim_fem_b200 import engine, meshgen
sys.path.insert(0, ".")
import bench
size = sys.argv[1] if len(sys.argv) > 1 else "L"
maxit = int(sys.argv[2]) if len(sys.argv) > 2 else 100
mesh = meshgen.synth_slab(size, contact_enabled=False)
confs = bench.sweep_definition(mesh, 8, 0)
ctx = engine.Context(0)
dm = ctx.mesh(mesh.nodes, mesh.tets, mesh.region, mesh.tris, mesh.bcid)
dm.assemble(bench.SIGMA); dm.bc_reset(8)
for k, c in enumerate(confs):
    dm.neumann_tris(c["tris"], bench.I_INJECT / c["area"], rhs=k)
dm.dirichlet(102, 0.0)
dm.solve(to_host=False, raise_on_noconv=False, maxit=maxit, check_every=50, use_graph=0, sample_spmv=4)
print(dm.last_stats)
