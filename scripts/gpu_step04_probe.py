import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "drivers"); sys.path.insert(0, "drivers/step03_ankle_layers"); sys.path.insert(0, "drivers/step04_pressure")
import tempfile
from pathlib import Path
import numpy as np
import pelvistim_fem_b200
from pelvistim_fem_b200 import engine, pipeline
import run_pressure_sweep as s4
p = s4.load_params()
levels = p["pressure_sweep"]["sigma_contact_Spm"]
d = Path(tempfile.mkdtemp())
mesh, e1, e2, bi = s4.build_mesh(p, d / "m")
print("mesh", mesh.nn, mesh.nt)
ctx = engine.Context(0)
dm = ctx.mesh(mesh.nodes, mesh.tets, mesh.region, mesh.tris, mesh.bcid); dm.pattern()
base = {1: 0.35, 2: 0.04, 3: 0.001}
for rep in range(2):
    t = time.perf_counter()
    dm.assemble([{**base, 4: s, 5: s} for s in levels]).bc_reset(1).neumann(101, 15.9).dirichlet(102, 0.0)
    dm.solve(to_host=False); ctx.sync()
    print("batched 15: %.1f ms" % ((time.perf_counter() - t) * 1e3), dm.last_stats)
    t = time.perf_counter(); its = []
    for s in levels:
        dm.assemble({**base, 4: s, 5: s}).bc_reset(1).neumann(101, 15.9).dirichlet(102, 0.0)
        dm.solve(to_host=False); its.append(dm.last_stats["iterations"])
    ctx.sync()
    print("sequential 15: %.1f ms" % ((time.perf_counter() - t) * 1e3), its)
