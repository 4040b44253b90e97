"""SpMV variant timings on the synthetic meshes (bench helper; prints GB/s against 12 nnz + 20 n bytes)."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import pelvistim_fem_b200 as pk
from pelvistim_fem_b200 import meshgen, engine
sizes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["M", "L"]
variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 3]
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 50
do_solve = len(sys.argv) > 4 and sys.argv[4] == "solve"
ctx = engine.Context(0)
sig = {1: 0.35, 2: 0.04, 3: 0.001, 4: 0.005, 5: 0.005}
out = {}
for size in sizes:
    m = meshgen.synth_slab(size)
    dm = ctx.mesh(m.nodes, m.tets, m.region, m.tris, m.bcid)
    nnz = dm.pattern()
    dm.assemble(sig); dm.bc_reset(1); dm.neumann(101, 15.975); dm.dirichlet(102, 0.0)
    bytes_ = 12 * nnz + 20 * m.nn
    print(size, "nn", m.nn, "nt", m.nt, "nnz", nnz, flush=True)
    # correctness of every variant against variant 1 on a random vector
    x = np.random.default_rng(0).standard_normal(m.nn)
    y1 = dm.spmv(x, 0, True, 1)
    for variant in variants:
        y = dm.spmv(x, 0, True, variant)
        err = np.abs(y - y1).max() / np.abs(y1).max()
        ms = dm.spmv_bench(variant, iters)
        print("   variant", variant, "err %.1e ms %.4f GB/s %.1f" % (err, ms, bytes_ / ms / 1e6), flush=True)
        out[f"{size}_v{variant}"] = dict(ms=ms, gbs=bytes_ / ms / 1e6, err=err)
    if do_solve:
        for variant in variants:
            dm.solve(to_host=False, spmv_variant=variant, rtol=1e-10)
            print("   solve variant", variant, dm.last_stats, flush=True)
            out[f"{size}_solve_v{variant}"] = dm.last_stats
    dm.close()
json.dump(out, open("gpurun_out/spmv.json", "w"), indent=1)
