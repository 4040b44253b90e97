"""Jacobi vs Chebyshev-polynomial preconditioning on one mesh: iterations, SpMV calls, solve time."""
import sys
sys.path.insert(0, ".")
import pelvistim_fem_b200
from pelvistim_fem_b200 import engine, meshgen
size = sys.argv[1] if len(sys.argv) > 1 else "L"
mesh = meshgen.synth_slab(size)
sig = {1: 0.35, 2: 0.04, 3: 0.001, 4: 0.005, 5: 0.005}
ctx = engine.Context(0)
dm = ctx.mesh(mesh.nodes, mesh.tets, mesh.region, mesh.tris, mesh.bcid)
dm.assemble(sig).bc_reset(1).neumann(101, 15.975).dirichlet(102, 0.0)
for kw in (dict(precond=0), dict(precond=1, cheb_degree=2), dict(precond=1, cheb_degree=3), dict(precond=1, cheb_degree=4),
           dict(precond=1, cheb_degree=4, cheb_ratio=60.0), dict(precond=1, cheb_degree=6, cheb_ratio=60.0)):
    for rep in range(2):
        dm.solve(to_host=False, rtol=1e-10, **kw)
    s = dm.last_stats
    print(size, kw, "iterations", s["iterations"], "spmv", s["spmv_calls"], "solve_ms %.1f" % s["solve_ms"], "true_rel %.1e" % s["true_rel_residual"], flush=True)
