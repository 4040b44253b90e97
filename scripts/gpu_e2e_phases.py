"""Wall-clock phases of one end-to-end sweep step (host buffers -> results) on the L mesh."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import pelvistim_fem_b200
from pelvistim_fem_b200 import engine, meshgen
import bench
size = sys.argv[1] if len(sys.argv) > 1 else "L"
mesh = meshgen.synth_slab(size, contact_enabled=False)
confs = bench.sweep_definition(mesh, 8, 0)
ctx = engine.Context(0)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
h = dict(nodes=pin(mesh.nodes), tets=pin(mesh.tets), region=pin(mesh.region), tris=pin(mesh.tris), bcid=pin(mesh.bcid))
phi_out = torch.empty((8, mesh.nn), dtype=torch.float64).pin_memory().numpy()
J_out = torch.empty((8, mesh.nn, 3), dtype=torch.float64).pin_memory().numpy()
for rep in range(3):
    T = {}
    def tic(name, t0):
        ctx.sync(); T[name] = time.perf_counter() - t0; return time.perf_counter()
    t = time.perf_counter()
    d = ctx.mesh(h["nodes"], h["tets"], h["region"], h["tris"], h["bcid"]); t = tic("mesh_create", t)
    d.pattern(); t = tic("pattern+geometry", t)
    d.assemble(bench.SIGMA); t = tic("assemble", t)
    d.bc_reset(8)
    for k, c in enumerate(confs):
        d.neumann_tris(c["tris"], bench.I_INJECT / c["area"], rhs=k)
    d.dirichlet(102, 0.0); t = tic("bc", t)
    d.solve(to_host=True, out=phi_out, rtol=bench.RTOL); t = tic("solve+d2h_phi", t)
    d.recover_current_batch(bench.RECOVER, to_host=True, out=J_out, wait=False); t = tic("recover_batch+d2h_J", t)
    reqs = []
    for k, c in enumerate(confs):
        fp = (c["center"][0], c["center"][1], c["r"], False)
        reqs += [dict(kind="nodes", sys=k, field=0, zmin=0.0397), dict(kind="nodes", sys=k, field=1, zmin=0.04 - 1e-5, mode=1, footprints=[fp]),
                 dict(kind="roi", sys=k, cen=[c["center"][0], c["center"][1], 0.03], r0=0.005, include_tris=False)]
    d.metrics_batch(reqs); t = tic("metrics_batch", t)
    t = time.perf_counter()
    d.close(); t = tic("close", t)
    print({k: round(v, 4) for k, v in T.items()}, "iters", d.last_stats["iterations"], "solve_ms", round(d.last_stats["solve_ms"], 1), flush=True)
