"""Restriction with several tasks per cell (split > 1: few, large cells): size M with the exactly inverted grid only
(coarse_levels=0 -> ~190 cells of ~2000 rows), 3 right-hand sides, fused and unfused kernels against the Jacobi solve."""
import os, sys
sys.path.insert(0, ".")
import numpy as np
import pelvistim_fem_b200  # noqa: F401
from pelvistim_fem_b200 import engine, meshgen

m = meshgen.synth_slab(sys.argv[1] if len(sys.argv) > 1 else "M")
SIG = {1: 0.35, 2: 0.04, 3: 0.001, 4: 0.005, 5: 0.005}
ref = None
for env in ({}, {"PTFEM_FUSE_UPDATE": "0"}, {"PTFEM_FUSE_PIPE": "2"}):
    for k in ("PTFEM_FUSE_UPDATE", "PTFEM_FUSE_PIPE"):
        os.environ.pop(k, None)
    os.environ.update(env)
    ctx = engine.Context(0)
    dm = ctx.mesh(m.nodes, m.tets, m.region, m.tris, m.bcid)
    dm.assemble(SIG).bc_reset(3)
    for k in range(3):
        dm.neumann(101, 10.0 + k, rhs=k)
    dm.dirichlet(102, 0.0)
    if ref is None:
        ref = dm.solve(precond=engine.PRECOND_JACOBI, rtol=1e-11).copy()
    phi = dm.solve(precond=engine.PRECOND_TWOLEVEL, coarse_nodes=300, coarse_levels=0, rtol=1e-11)
    st = dm.last_stats
    err = max(np.abs(phi[k] - ref[k]).max() / np.abs(ref[k]).max() for k in range(3))
    print(env, "iterations", st["iterations"], "converged", st["converged"], "rel_err_vs_jacobi", f"{err:.2e}", flush=True)
    assert err < 1e-7, err
    dm.close(); ctx.close()
print("ok")
