"""One multi-RHS solve of the bench workload with a chosen preconditioner (profiling target)."""
import os, sys
sys.path.insert(0, ".")
import pelvistim_fem_b200
from pelvistim_fem_b200 import engine, meshgen
import bench

size = sys.argv[1] if len(sys.argv) > 1 else "L"
nrhs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
kw = dict(precond=int(sys.argv[3]) if len(sys.argv) > 3 else 2)
if len(sys.argv) > 4:
    kw["coarse_nodes"] = int(sys.argv[4])
if len(sys.argv) > 5:
    kw["coarse_levels"] = int(sys.argv[5])
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 1
mesh = meshgen.synth_slab(size, contact_enabled=False)
confs = bench.sweep_definition(mesh, 8)
ctx = engine.Context(0)
dm = ctx.mesh(mesh.nodes, mesh.tets, mesh.region, mesh.tris, mesh.bcid)
for rep in range(reps):
    dm.assemble(bench.SIGMA)
    dm.bc_reset(nrhs)
    for k in range(nrhs):
        dm.neumann_tris(confs[k]["tris"], bench.I_INJECT / confs[k]["area"], rhs=k)
    dm.dirichlet(102, 0.0)
    dm.solve(to_host=False, rtol=1e-10, use_graph=int(os.environ.get("GRAPH", "0")), **kw)
    s = dm.last_stats
    print(size, nrhs, kw, "its", s["iterations"], "solve_ms %.1f" % s["solve_ms"], "setup_ms %.1f" % s["setup_ms"], flush=True)
