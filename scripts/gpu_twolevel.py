"""Two-level / BPX preconditioner against Jacobi on the bench workload: iterations, set-up and solve time, agreement."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import pelvistim_fem_b200
from pelvistim_fem_b200 import engine, meshgen
import bench

sizes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["S", "M", "L"]
variants = [dict(precond=0), dict(precond=2), dict(precond=2, coarse_levels=0), dict(precond=2, coarse_levels=1), dict(precond=2, coarse_nodes=1000)]
ctx = engine.Context(0)
for size in sizes:
    mesh = meshgen.synth_slab(size, contact_enabled=False)
    confs = bench.sweep_definition(mesh, 8)
    dm = ctx.mesh(mesh.nodes, mesh.tets, mesh.region, mesh.tris, mesh.bcid)
    ref = None
    for nrhs in (8, 1):
        for kw in variants:
            dm.assemble(bench.SIGMA)
            dm.bc_reset(nrhs)
            for k in range(nrhs):
                dm.neumann_tris(confs[k]["tris"], bench.I_INJECT / confs[k]["area"], rhs=k)
            dm.dirichlet(102, 0.0)
            try:
                t0 = time.perf_counter()
                phi = dm.solve(to_host=(size != "L"), rtol=1e-10, **kw)
                wall = (time.perf_counter() - t0) * 1e3
            except Exception as e:  # noqa: BLE001
                print(size, nrhs, kw, "FAILED", e, flush=True)
                continue
            s = dm.last_stats
            err = ""
            if phi is not None:
                if kw == dict(precond=0):
                    ref = phi.copy()
                elif ref is not None and ref.shape == phi.shape:
                    err = " maxrel %.2e" % (np.abs(phi - ref).max() / np.abs(ref).max())
            print(size, "nrhs", nrhs, kw, "its", s["iterations"], "solve_ms %.1f" % s["solve_ms"], "setup_ms %.1f" % s["setup_ms"],
                  "wall_ms %.1f" % wall, "k", s["coarse_unknowns"], "true_rel %.1e" % s["true_rel_residual"], err, flush=True)
        ref = None
    dm.close() if hasattr(dm, "close") else None
