"""Wall/device time of the phases of one bench sweep step (size L by default), each bracketed by a context sync:
    python scripts/gpu_step_phases.py [size] [reps]
Phases: assemble, bc (reset + 8 Neumann patches + Dirichlet), solve (incl. BC elimination + preconditioner set-up),
recovery x nconf, metrics x nconf.  Prints one JSON line (median over reps)."""
import json, statistics, sys, time
sys.path.insert(0, ".")
import numpy as np
import bench
import pelvistim_fem_b200  # noqa: F401
from pelvistim_fem_b200 import engine, meshgen

size = sys.argv[1] if len(sys.argv) > 1 else "L"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
mesh = meshgen.synth_slab(size, contact_enabled=False)
confs = bench.sweep_definition(mesh, 8, 0)
ctx = engine.Context(0)
dm = ctx.mesh(mesh.nodes, mesh.tets, mesh.region, mesh.tris, mesh.bcid)
dm.pattern()
Lz, t_skin = mesh.meta["Lz"], mesh.meta["t_skin"]
acc = {}


def tick(name, t0):
    ctx.sync()
    acc.setdefault(name, []).append((time.perf_counter() - t0) * 1e3)
    return time.perf_counter()


for rep in range(reps + 2):
    if rep == 2:
        acc.clear()
    ctx.sync()
    t_all = t0 = time.perf_counter()
    dm.assemble(bench.step_sigma(rep))
    t0 = tick("assemble", t0)
    dm.bc_reset(len(confs))
    for k, c in enumerate(confs):
        dm.neumann_tris(c["tris"], bench.I_INJECT / c["area"], rhs=k)
    dm.dirichlet(102, 0.0)
    t0 = tick("bc", t0)
    dm.solve(to_host=False, rtol=bench.RTOL, precond=-1)
    st = dm.last_stats
    t0 = tick("solve_call", t0)
    acc.setdefault("solve_ms_stat", []).append(st["solve_ms"])
    acc.setdefault("setup_ms_stat", []).append(st["setup_ms"])
    acc.setdefault("iterations", []).append(st["iterations"])
    trec = tmet = 0.0
    for k, c in enumerate(confs):
        t1 = time.perf_counter()
        dm.recover_current(k, "lumped", to_host=False)
        ctx.sync()
        t2 = time.perf_counter()
        fp = (c["center"][0], c["center"][1], c["r"], False)
        dm.metric_nodes(0, Lz - 0.2 * t_skin, sys=k)
        dm.metric_nodes(1, Lz - 1e-5, mode=1, footprints=[fp], scale_r=1.0, sys=k)
        dm.metric_roi([c["center"][0], c["center"][1], Lz - 0.010], 0.005, (1.0, 1.5, 2.0, 3.0), include_tris=False, sys=k)
        ctx.sync()
        t3 = time.perf_counter()
        trec += (t2 - t1) * 1e3
        tmet += (t3 - t2) * 1e3
    acc.setdefault("recover_x8", []).append(trec)
    acc.setdefault("metrics_x8", []).append(tmet)
    acc.setdefault("step_total", []).append((time.perf_counter() - t_all) * 1e3)
print(json.dumps({k: statistics.median(v) for k, v in acc.items()}))
