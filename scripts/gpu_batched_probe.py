"""Batched-values PCG (K9): the 15 sigma_contact levels of step04 as 15 matrices on one pattern, on a synthetic mesh."""
import sys, time
sys.path.insert(0, ".")
import yaml
import pelvistim_fem_b200
from pelvistim_fem_b200 import engine, meshgen
size = sys.argv[1] if len(sys.argv) > 1 else "L"
levels = yaml.safe_load(open("tests/golden/step04_params.yaml"))["pressure_sweep"]["sigma_contact_Spm"]
mesh = meshgen.synth_slab(size)
ctx = engine.Context(0)
dm = ctx.mesh(mesh.nodes, mesh.tets, mesh.region, mesh.tris, mesh.bcid)
nnz = dm.pattern()
base = {1: 0.35, 2: 0.04, 3: 0.001}
t = time.perf_counter()
dm.assemble([{**base, 4: s, 5: s} for s in levels]).bc_reset(1).neumann(101, 15.975).dirichlet(102, 0.0)
ctx.sync(); ta = time.perf_counter() - t
dm.solve(to_host=False, rtol=1e-10, sample_spmv=4)
s = dm.last_stats
by = (8 * 16 + 4) * nnz + 4 * mesh.nn + 16 * 16 * mesh.nn
print(f"{size}: 15 matrices batched: assemble+BC {ta*1e3:.1f} ms, {s['iterations']} iterations, solve {s['solve_ms']:.1f} ms "
      f"({s['solve_ms']/s['iterations']:.3f} ms/it), batched SpMV {s['spmv_ms']:.3f} ms = {by/s['spmv_ms']/1e6:.0f} GB/s "
      f"({by/s['spmv_ms']/1e6/6543.7:.2f} of peak), true_rel {s['true_rel_residual']:.1e}", flush=True)
t = time.perf_counter(); tot_it = 0
for sc in levels[:3]:
    dm.assemble({**base, 4: sc, 5: sc}).bc_reset(1).neumann(101, 15.975).dirichlet(102, 0.0)
    dm.solve(to_host=False, rtol=1e-10); tot_it += dm.last_stats["iterations"]
ctx.sync()
print(f"   one by one (first 3 levels): {(time.perf_counter()-t)/3*1e3:.1f} ms per level, {tot_it/3:.0f} iterations each")
