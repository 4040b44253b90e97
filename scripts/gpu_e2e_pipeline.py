"""Host wall-clock of the calls of bench.py's pipelined end-to-end step (two meshes alive, outputs double-buffered)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import pelvistim_fem_b200  # noqa
from pelvistim_fem_b200 import engine, meshgen
import bench
size = sys.argv[1] if len(sys.argv) > 1 else "L"
mesh = meshgen.synth_slab(size, contact_enabled=False)
confs = bench.sweep_definition(mesh, 8, 0)
ctx = engine.Context(0)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
h = dict(nodes=pin(mesh.nodes), tets=pin(mesh.tets), region=pin(mesh.region), tris=pin(mesh.tris), bcid=pin(mesh.bcid))
outs = [(torch.empty((8, mesh.nn), dtype=torch.float64).pin_memory().numpy(), torch.empty((8, mesh.nn, 3), dtype=torch.float64).pin_memory().numpy()) for _ in range(2)]
prev = None
for s in range(8):
    T = {}
    t0 = time.perf_counter()
    d = ctx.mesh(h["nodes"], h["tets"], h["region"], h["tris"], h["bcid"]); T["mesh"] = time.perf_counter() - t0; t = time.perf_counter()
    d.pattern(); T["pattern"] = time.perf_counter() - t; t = time.perf_counter()
    if prev is not None:
        prev.close()
    T["close_prev"] = time.perf_counter() - t; t = time.perf_counter()
    po, jo = outs[s % 2]
    bench.run_sweep_step(d, mesh, confs, s, phi_out=po, J_out=jo); T["sweep"] = time.perf_counter() - t
    prev = d
    T["total"] = time.perf_counter() - t0
    print({k: round(v * 1e3, 2) for k, v in T.items()}, flush=True)
prev.close()
