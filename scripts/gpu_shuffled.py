"""SpMV/SpMM timings on a mesh whose node numbering was shuffled (worst-case locality), Morton order off/on."""
import sys, os
sys.path.insert(0, ".")
import numpy as np
import pelvistim_fem_b200
from pelvistim_fem_b200 import engine, meshgen
import bench
size = sys.argv[1] if len(sys.argv) > 1 else "M"
mesh = meshgen.synth_slab(size, contact_enabled=False)
confs = bench.sweep_definition(mesh, 8, 0)
rng = np.random.default_rng(1)
for mode in ("natural", "block-shuffled", "random"):
    if mode == "natural":
        perm = np.arange(mesh.nn)
    elif mode == "random":
        perm = rng.permutation(mesh.nn)
    else:   # blocks of 512 consecutive nodes in random order (coherent pieces, like a mesher's entity order)
        nb = (mesh.nn + 511) // 512
        perm = np.concatenate([np.arange(b * 512, min(mesh.nn, (b + 1) * 512)) for b in rng.permutation(nb)])
    inv = np.empty_like(perm); inv[perm] = np.arange(mesh.nn)
    nodes, tets, tris = mesh.nodes[perm], inv[mesh.tets].astype(np.int32), inv[mesh.tris].astype(np.int32)
    for morton in (0, 1, -1):
        os.environ["PTFEM_MORTON"] = str(morton)
        ctx = engine.Context(0)
        dm = ctx.mesh(nodes, tets, mesh.region, tris, mesh.bcid)
        nnz = dm.pattern()
        dm.assemble(bench.SIGMA); dm.bc_reset(8)
        for k, cf in enumerate(confs):
            dm.neumann_tris(cf["tris"], bench.I_INJECT / cf["area"], rhs=k)
        dm.dirichlet(102, 0.0)
        dm.solve(to_host=False, raise_on_noconv=False, maxit=100, check_every=50, sample_spmv=8, spmv_variant=2)
        spmm = dm.last_stats["spmv_ms"]
        dm.bc_reset(1); dm.neumann_tris(confs[0]["tris"], 1.0); dm.dirichlet(102, 0.0)
        s1 = dm.spmv_bench(2, 30); v1 = dm.spmv_bench(1, 30)
        print(f"{size} {mode:15s} morton={morton}  spmm8 {spmm:.4f} ms  spmv1 stream {s1:.4f} ms ({(12*nnz+20*mesh.nn)/s1/1e6:.0f} GB/s)  vector {v1:.4f} ms", flush=True)
        dm.close(); ctx.close()
