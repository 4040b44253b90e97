"""8-RHS solve of the bench sweep (size L by default) under tuning-knob combinations; one JSON line per combination.
    python scripts/gpu_pcg_knobs.py [size] KNOB=v1,v2 [KNOB2=...]        e.g.  PTFEM_PUPDATE_NP=1,2 PTFEM_RESTRICT_OCC=0,1
Every combination gets a fresh context (the knobs are read from the environment when the context is created)."""
import itertools, json, os, sys
sys.path.insert(0, ".")
import bench
import pelvistim_fem_b200  # noqa: F401
from pelvistim_fem_b200 import engine, meshgen

args = [a for a in sys.argv[1:]]
size = args.pop(0) if args and "=" not in args[0] else "L"
knobs = [(a.split("=")[0], a.split("=")[1].split(",")) for a in args]
mesh = meshgen.synth_slab(size, contact_enabled=False)
confs = bench.sweep_definition(mesh, 8, 0)
for combo in itertools.product(*[v for _, v in knobs]) if knobs else [()]:
    for (k, _), v in zip(knobs, combo):
        os.environ[k] = v
    ctx = engine.Context(0)
    dm = ctx.mesh(mesh.nodes, mesh.tets, mesh.region, mesh.tris, mesh.bcid)
    dm.pattern()
    best = None
    for rep in range(4):
        dm.assemble(bench.step_sigma(rep))
        dm.bc_reset(len(confs))
        for k, c in enumerate(confs):
            dm.neumann_tris(c["tris"], bench.I_INJECT / c["area"], rhs=k)
        dm.dirichlet(102, 0.0)
        dm.solve(to_host=False, rtol=bench.RTOL, precond=-1)
        st = dm.last_stats
        if rep >= 1 and (best is None or st["solve_ms"] < best["solve_ms"]):
            best = dict(st)
    print(json.dumps(dict(knobs=dict(zip([k for k, _ in knobs], combo)), solve_ms=best["solve_ms"], setup_ms=best["setup_ms"],
                          iterations=best["iterations"], ms_per_iteration=best["solve_ms"] / max(best["iterations"], 1),
                          coarse_unknowns=best["coarse_unknowns"])), flush=True)
    dm.close()
    ctx.close()
