"""Env-knob sweep for the multi-RHS PCG kernels on one mesh: reports spmv_ms (S=8) and 1-RHS SpMV."""
import sys, os, json
sys.path.insert(0, ".")
import numpy as np
import pelvistim_fem_b200
from pelvistim_fem_b200 import engine, meshgen
import bench
size = sys.argv[1] if len(sys.argv) > 1 else "L"
mesh = meshgen.synth_slab(size, contact_enabled=False)
confs = bench.sweep_definition(mesh, 8, 0)
configs = [dict(MORTON=0), dict(MORTON=1), dict(MORTON=1, STREAM_ROWS=32), dict(MORTON=1, STREAM_ROWS=128), dict(MORTON=1, INTERLEAVE=0)]
if len(sys.argv) > 2:
    configs = [eval("dict(" + a + ")") for a in sys.argv[2:]]
for c in configs:
    for k in ("XPREFETCH", "STREAM_ROWS", "INTERLEAVE", "STREAM_STAGES", "MORTON", "STREAM_CAP", "CTAS_PER_SM", "SPMM_WINDOW", "WINDOW_BX", "WINDOW_CTAS", "FUSE_UPDATE"):
        os.environ.pop("PTFEM_" + k, None)
    for k, v in c.items():
        os.environ["PTFEM_" + k] = str(v)
    ctx = engine.Context(0)
    dm = ctx.mesh(mesh.nodes, mesh.tets, mesh.region, mesh.tris, mesh.bcid)
    nnz = dm.pattern()
    print('window plan', dm.window_plan(), flush=True)
    dm.assemble(bench.SIGMA); dm.bc_reset(8)
    for k, cf in enumerate(confs):
        dm.neumann_tris(cf["tris"], bench.I_INJECT / cf["area"], rhs=k)
    dm.dirichlet(102, 0.0)
    out = {}
    for variant in (2, 1):
        dm.solve(to_host=False, raise_on_noconv=False, maxit=200, check_every=50, sample_spmv=8, spmv_variant=variant)
        st = dm.last_stats
        out[f"v{variant}_spmm_ms"] = round(st["spmv_ms"], 4)
        out[f"v{variant}_ms_per_it"] = round(st["solve_ms"] / st["iterations"], 4)
    dm.bc_reset(1); dm.neumann_tris(confs[0]["tris"], 1.0); dm.dirichlet(102, 0.0)
    out["spmv1_ms"] = round(dm.spmv_bench(2, 30), 4)
    out["spmv1_gbs"] = round((12 * nnz + 20 * mesh.nn) / out["spmv1_ms"] / 1e6, 1)
    out["spmm_gbs"] = round((12 * nnz + 4 * mesh.nn + 128 * mesh.nn) / out["v2_spmm_ms"] / 1e6, 1)
    print(c, out, flush=True)
    dm.close(); ctx.close()
