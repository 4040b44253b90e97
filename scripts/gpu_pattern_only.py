"""Mesh upload + pattern/geometry (+ window plan) only, for a launch list of the once-per-mesh kernels."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import pelvistim_fem_b200
from pelvistim_fem_b200 import engine, meshgen
size = sys.argv[1] if len(sys.argv) > 1 else "L"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
mesh = meshgen.synth_slab(size, contact_enabled=False)
ctx = engine.Context(0)
for rep in range(reps):
    t0 = time.perf_counter()
    d = ctx.mesh(mesh.nodes, mesh.tets, mesh.region, mesh.tris, mesh.bcid); ctx.sync(); t1 = time.perf_counter()
    d.pattern(); ctx.sync(); t2 = time.perf_counter()
    print("mesh_create %.1f ms  pattern %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3), d.window_plan(), flush=True)
    d.close()
