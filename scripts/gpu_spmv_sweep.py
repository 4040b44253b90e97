"""Sweep the SpMV tuning knobs (env-driven) on one synthetic mesh; one process, one mesh generation."""
import sys, os, json, itertools
import numpy as np
sys.path.insert(0, ".")
import pelvistim_fem_b200 as pk
from pelvistim_fem_b200 import meshgen, engine
size = sys.argv[1] if len(sys.argv) > 1 else "L"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
m = meshgen.synth_slab(size)
sig = {1: 0.35, 2: 0.04, 3: 0.001, 4: 0.005, 5: 0.005}
x = np.random.default_rng(0).standard_normal(m.nn)
res = []
y_ref = None
configs = [dict(v=1)]
for (rows, cap) in ((128, 2048), (128, 2304), (96, 1536), (64, 1024), (32, 512)):
    for st in (2, 3, 4):
        for tpr in (1, 2):
            configs.append(dict(v=2, ROWS=rows, CAP=cap, ST=st, TPR=tpr))
for c in configs:
    os.environ["PTFEM_INTERLEAVE"] = str(c.get("IL", 1))
    os.environ["PTFEM_STREAM_STAGES"] = str(c.get("ST", 2))
    os.environ["PTFEM_STREAM_TPR"] = str(c.get("TPR", 4))
    os.environ["PTFEM_CTAS_PER_SM"] = str(c.get("CPS", 0))
    os.environ["PTFEM_STREAM_ROWS"] = str(c.get("ROWS", 0))
    os.environ["PTFEM_STREAM_CAP"] = str(c.get("CAP", 0))
    ctx = engine.Context(0)
    dm = ctx.mesh(m.nodes, m.tets, m.region, m.tris, m.bcid)
    nnz = dm.pattern()
    dm.assemble(sig); dm.bc_reset(1); dm.neumann(101, 15.975); dm.dirichlet(102, 0.0)
    bytes_ = 12 * nnz + 20 * m.nn
    y = dm.spmv(x, 0, True, c["v"])
    if y_ref is None:
        y_ref = y
    err = float(np.abs(y - y_ref).max() / np.abs(y_ref).max())
    ms = dm.spmv_bench(c["v"], iters)
    ms2 = dm.spmv_bench(c["v"], iters)
    r = dict(c, ms=min(ms, ms2), gbs=bytes_ / min(ms, ms2) / 1e6, err=err)
    print(r, flush=True)
    res.append(r)
    dm.close(); ctx.close()
json.dump(res, open("gpurun_out/spmv_sweep.json", "w"), indent=1)
