"""First GPU bring-up: parity vs oracle on small meshes, SpMV timings on large ones."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import pelvistim_fem_b200 as pk
from pelvistim_fem_b200 import meshgen, engine
from oracle import fem_oracle as fo

ctx = engine.Context(0)
out = {}
sig = {1: 0.35, 2: 0.04, 3: 0.001, 4: 0.005, 5: 0.005}
for size in ["XS", "S"]:
    m = meshgen.synth_slab(size)
    ref = fo.solve_case(m, sig, [(102, 0.0)], [(101, 15.975)], recover="l2")
    rp, col = fo.csr_pattern(m.nn, m.tets)
    for variant in (1, 2, 3):
        res = engine.solve_case(ctx, m, sig, [(102, 0.0)], [(101, 15.975)], recover="l2", spmv_variant=variant, rtol=1e-12)
        dm = res["dmesh"]
        grp, gcol = dm.get_pattern()
        val = dm.get_values(0, False)
        Kraw = ref["K_raw"]
        e_phi = np.abs(res["phi"] - ref["phi"]).max() / np.abs(ref["phi"]).max()
        e_J = np.abs(res["J"] - ref["J"]).max() / np.abs(ref["J"]).max()
        print(size, "variant", variant, "pattern", np.array_equal(grp, rp) and np.array_equal(gcol, col),
              "val err", np.abs(val - Kraw.data).max() / np.abs(Kraw.data).max(),
              "phi err %.2e J err %.2e" % (e_phi, e_J), res["stats"], flush=True)
        x = np.random.default_rng(0).standard_normal(m.nn)
        y = dm.spmv(x, 0, False, variant)
        print("   spmv err", np.abs(y - Kraw @ x).max() / np.abs(Kraw @ x).max(), flush=True)
        dm.close()
    # chebyshev
    res = engine.solve_case(ctx, m, sig, [(102, 0.0)], [(101, 15.975)], recover="l2", precond=1, cheb_degree=4)
    print(size, "cheb phi err %.2e" % (np.abs(res["phi"] - ref["phi"]).max() / np.abs(ref["phi"]).max()), res["stats"], flush=True)
    res["dmesh"].close()

for size in ["M", "L"]:
    t = time.time(); m = meshgen.synth_slab(size); tm = time.time() - t
    t = time.time()
    dm = ctx.mesh(m.nodes, m.tets, m.region, m.tris, m.bcid)
    nnz = dm.pattern(); ctx.sync(); tp = time.time() - t
    t = time.time(); dm.assemble(sig); dm.bc_reset(1); dm.neumann(101, 15.975); dm.dirichlet(102, 0.0); ctx.sync(); ta = time.time() - t
    bytes_ = 12 * nnz + 20 * m.nn
    print(size, "nn", m.nn, "nt", m.nt, "nnz", nnz, "mesh %.1fs pattern %.2fs assemble %.3fs" % (tm, tp, ta), flush=True)
    for variant in (1, 2, 3):
        ms = dm.spmv_bench(variant, 50)
        print("   variant", variant, "ms %.4f GB/s %.1f" % (ms, bytes_ / ms / 1e6), flush=True)
        out[f"{size}_v{variant}"] = dict(ms=ms, gbs=bytes_ / ms / 1e6)
    for variant in (1, 2):
        t = time.time()
        dm.solve(to_host=False, spmv_variant=variant, rtol=1e-10)
        ctx.sync()
        print("   solve variant", variant, "wall %.3f" % (time.time() - t), dm.last_stats, flush=True)
        out[f"{size}_solve_v{variant}"] = dm.last_stats
    dm.close()
json.dump(out, open("gpurun_out/first.json", "w"), indent=1)
