"""Window SpMM on the 67 M-tet slab (11.3 M rows): plan, an 8-RHS solve with it against the same solve on the streaming kernel."""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np
import pelvistim_fem_b200
from pelvistim_fem_b200 import engine, meshgen
dims = tuple(int(v) for v in sys.argv[1].split("x")) if len(sys.argv) > 1 else (288, 216, 180)
mesh = meshgen.synth_slab(dims, contact_enabled=False)
print("mesh", mesh.nn, mesh.nt, flush=True)
out = {}
for flag in ("1", "0"):
    os.environ["PTFEM_SPMM_WINDOW"] = flag
    ctx = engine.Context(0)
    dm = ctx.mesh(mesh.nodes, mesh.tets, mesh.region, mesh.tris, mesh.bcid)
    nnz = dm.pattern()
    print("window", flag, dm.window_plan(), flush=True)
    dm.assemble({1: 0.35, 2: 0.04, 3: 0.001}).bc_reset(8)
    for k in range(8):
        dm.neumann(101, 15.975 * (1 + 0.1 * k), rhs=k)
    dm.dirichlet(102, 0.0)
    phi = dm.solve(rtol=1e-10, sample_spmv=8)
    s = dm.last_stats
    by = 12 * nnz + 4 * mesh.nn + 128 * mesh.nn
    print("  %d iterations, solve %.1f ms, SpMM %.4f ms = %.0f GB/s (%.2f of measured peak), true rel. residual %.1e"
          % (s["iterations"], s["solve_ms"], s["spmv_ms"], by / s["spmv_ms"] / 1e6, by / s["spmv_ms"] / 1e6 / 6543.7, s["true_rel_residual"]), flush=True)
    out[flag] = phi[[0, 7]].copy()
    dm.close(); ctx.close()
print("max rel. difference window vs streaming: %.2e" % (np.abs(out["1"] - out["0"]).max() / np.abs(out["0"]).max()))
