"""Small end-to-end exercise of every kernel family (for compute-sanitizer runs)."""
import sys
sys.path.insert(0, ".")
import numpy as np
import pelvistim_fem_b200
from pelvistim_fem_b200 import engine, meshgen, pipeline
ctx = engine.Context(0)
sig = {1: 0.35, 2: 0.04, 3: 0.001, 4: 0.005, 5: 0.005}
m = meshgen.synth_slab("XS")
for variant in (1, 2, 3):
    r = engine.solve_case(ctx, m, sig, [(102, 0.0)], [(101, 15.975)], recover="l2", spmv_variant=variant)
    dm = r["dmesh"]
    dm.metric_nodes(0, 0.039); dm.metric_roi([0.015, 0.045, 0.03], 0.005); dm.metric_jstats(0.0)
    dm.metric_pad_current(0.0404, (0.015, 0.045, 0.01, False)); dm.metric_column_fit(0.04, 0.03, 0.01); dm.metric_reaction(102)
    dm.sample_polyline(np.stack([np.linspace(0.01, 0.07, 50), np.full(50, 0.03), np.full(50, 0.02)], axis=1))
    dm.recover_current(0, "lumped"); dm.recover_current(0, "average"); dm.element_fields(0)
    dm.close()
dm = ctx.mesh(m.nodes, m.tets, m.region, m.tris, m.bcid)
dm.assemble([{**sig, 4: s, 5: s} for s in (5e-5, 5e-3, 0.5)]).bc_reset(1).neumann(101, 15.975).dirichlet(102, 0.0)
dm.solve(); dm.solve(precond=1, cheb_degree=3)
dm.assemble(sig).bc_reset(5)
top = np.nonzero(m.bcid == 101)[0].astype(np.int32)
for k in range(5):
    dm.neumann_tris(top, 10.0 + k, rhs=k)
dm.dirichlet(102, 0.0)
dm.solve(spmv_variant=2); dm.solve(spmv_variant=1, precond=1)
dm.set_coords(m.nodes * np.array([1.0, 1.0, 0.95]))
dm.assemble(sig).bc_reset(1).neumann(101, 1.0).dirichlet(102, 0.0)
dm.solve()
dm.close()
ms = meshgen.box_mesh(0.04, 0.04, 0.02, 6, 4, 2, jitter=0.2, seed=7, ids=(2, 1, 3))
engine.solve_case(ctx, ms, {1: 0.2}, [(2, 1.0), (1, 0.0)], [], recover="l2")["dmesh"].close()
print("sanity_small ok")
