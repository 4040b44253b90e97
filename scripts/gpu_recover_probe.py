"""One 8-RHS solve (few iterations) + one nodal current recovery + metrics on the L mesh (for ncu launch lists)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import pelvistim_fem_b200
from pelvistim_fem_b200 import engine, meshgen
import bench
size = sys.argv[1] if len(sys.argv) > 1 else "L"
mesh = meshgen.synth_slab(size, contact_enabled=False)
confs = bench.sweep_definition(mesh, 8, 0)
ctx = engine.Context(0)
dm = ctx.mesh(mesh.nodes, mesh.tets, mesh.region, mesh.tris, mesh.bcid)
dm.assemble(bench.SIGMA); dm.bc_reset(8)
for k, c in enumerate(confs):
    dm.neumann_tris(c["tris"], bench.I_INJECT / c["area"], rhs=k)
dm.dirichlet(102, 0.0)
dm.solve(to_host=False, raise_on_noconv=False, maxit=50, check_every=50, use_graph=0)
ctx.sync()
for rep in range(2):
    t = time.perf_counter()
    dm.recover_current(1, "l2", to_host=False); ctx.sync()
    t1 = time.perf_counter() - t
    t = time.perf_counter()
    c = confs[1]
    dm.metric_nodes(0, 0.0397, sys=1); dm.metric_nodes(1, 0.04 - 1e-5, mode=1, footprints=[(c["center"][0], c["center"][1], c["r"], False)], sys=1)
    dm.metric_roi([c["center"][0], c["center"][1], 0.03], 0.005, (1.0, 1.5, 2.0, 3.0), include_tris=False, sys=1)
    ctx.sync()
    print("recover %.1f ms, metrics %.1f ms" % (t1 * 1e3, (time.perf_counter() - t) * 1e3), flush=True)
