"""One rank of the row-partitioned solve with coarse grids through the peer-memory kernels, no check solves
(for `ncu --metrics gpu__time_duration.sum -k regex:...` launch lists of the S = 1 iteration).
    python scripts/gpu_dist_coarse_probe.py L [p2p|single]"""
import json, os, sys
sys.path.insert(0, ".")
import torch
import torch.distributed as dist
import pelvistim_fem_b200
from pelvistim_fem_b200 import engine, meshgen, distsolve

size = sys.argv[1] if len(sys.argv) > 1 else "L"
transport = sys.argv[2] if len(sys.argv) > 2 else "p2p"
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29513")
dist.init_process_group("gloo", rank=0, world_size=1)
mesh = meshgen.synth_slab(size)
ctx = engine.Context(0)
res = distsolve.partitioned_solve(ctx, mesh, {1: 0.35, 2: 0.04, 3: 0.001, 4: 0.005, 5: 0.005}, [(102, 0.0)], [(101, 15.975)], 0, 1,
                                  check=False, transport=transport, force_p2p=transport == "p2p", rtol=1e-10)
print(json.dumps(dict(size=size, transport=res["transport"], coarse=res["coarse"], iterations=res["stats"]["iterations"],
                      solve_ms=res["stats"]["solve_ms"], **res["timings"])), flush=True)
dist.destroy_process_group()
