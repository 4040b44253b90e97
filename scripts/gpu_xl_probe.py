"""Largest-mesh probe: SpMV / 1-RHS PCG iteration on a synthetic slab much larger than L (vector > L2)."""
import sys, time, resource
sys.path.insert(0, ".")
import numpy as np
import pelvistim_fem_b200
from pelvistim_fem_b200 import engine, meshgen
dims = tuple(int(v) for v in sys.argv[1].split("x")) if len(sys.argv) > 1 else (288, 216, 180)
t = time.time(); mesh = meshgen.synth_slab(dims, contact_enabled=False); print("mesh", mesh.nn, mesh.nt, "gen %.1fs" % (time.time() - t), flush=True)
ctx = engine.Context(0)
t = time.time(); dm = ctx.mesh(mesh.nodes, mesh.tets, mesh.region, mesh.tris, mesh.bcid); nnz = dm.pattern(); ctx.sync()
print("upload+pattern %.2fs nnz %d" % (time.time() - t, nnz), flush=True)
dm.assemble({1: 0.35, 2: 0.04, 3: 0.001}).bc_reset(1).neumann(101, 15.975).dirichlet(102, 0.0)
by = 12 * nnz + 20 * mesh.nn
for v in (2, 1):
    ms = dm.spmv_bench(v, 20)
    print("variant", v, "%.4f ms  %.0f GB/s (%.2f of measured peak)" % (ms, by / ms / 1e6, by / ms / 1e6 / 6543.7), flush=True)
dm.solve(to_host=False, raise_on_noconv=False, maxit=200, check_every=50, rtol=1e-10, precond=engine.PRECOND_JACOBI)
s = dm.last_stats
print("200 Jacobi-PCG iterations: %.1f ms (%.3f ms/it)" % (s["solve_ms"], s["solve_ms"] / 200), "host maxrss GB %.1f" % (resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e6), flush=True)
# full solves to rtol 1e-10: default (Jacobi + coarse grids), then plain Jacobi
for name, pc in (("auto", engine.PRECOND_AUTO), ("auto again (set-up reused)", engine.PRECOND_AUTO), ("jacobi", engine.PRECOND_JACOBI)):
    dm.solve(to_host=False, rtol=1e-10, precond=pc)
    s = dm.last_stats
    print("full solve %s: precond %d, %d iterations, solve %.1f ms, set-up %.1f ms, coarse unknowns %d, true rel. residual %.1e"
          % (name, s["precond"], s["iterations"], s["solve_ms"], s["setup_ms"], s["coarse_unknowns"], s["true_rel_residual"]), flush=True)
import torch
print("device memory in use GB %.1f" % ((torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 1e9))
