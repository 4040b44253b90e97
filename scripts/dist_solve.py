"""Row-partitioned solve of one synthetic mesh over the ranks of a torchrun launch (or 1 rank).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_solve.py M
    ... scripts/dist_solve.py L p2p jacobi          # transport, preconditioner (auto = Jacobi + coarse grids | jacobi)
    ... scripts/dist_solve.py XL p2p auto dist      # distributed set-up: no rank uploads the whole mesh (owned nodes + ghost elements)
    PTFEM_SAME_GPU=1 ... --nproc-per-node 2 scripts/dist_solve.py M p2p   # all ranks on GPU 0 (CUDA IPC between processes of
                                                    # one device; time-sliced, slow - a protocol check when one GPU is all there is)
"""
import json, os, sys
sys.path.insert(0, ".")
import numpy as np
import torch
import torch.distributed as dist
import pelvistim_fem_b200
from pelvistim_fem_b200 import engine, meshgen, distsolve

size = sys.argv[1] if len(sys.argv) > 1 else "M"
transport = sys.argv[2] if len(sys.argv) > 2 else "p2p"
precond = sys.argv[3] if len(sys.argv) > 3 else "auto"
distributed = len(sys.argv) > 4 and sys.argv[4] == "dist"
same_gpu = os.environ.get("PTFEM_SAME_GPU", "0") == "1"
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if same_gpu:
    local = 0
torch.cuda.set_device(local)
if world > 1 and not same_gpu:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
else:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29512")
    dist.init_process_group("gloo", rank=rank, world_size=world)
mesh = meshgen.synth_slab(size)
sig = {1: 0.35, 2: 0.04, 3: 0.001, 4: 0.005, 5: 0.005}
ctx = engine.Context(local)
res = distsolve.partitioned_solve(ctx, mesh, sig, [(102, 0.0)], [(101, 15.975)], rank, world, check=not distributed, transport=transport,
                                  coarse=precond == "auto", force_p2p=os.environ.get("PTFEM_FORCE_P2P", "0") == "1", distributed=distributed,
                                  rtol=1e-10)
line = dict(rank=rank, world=world, size=size, transport=res["transport"], coarse=res["coarse"], coarse_note=res["coarse_note"],
            nloc=res["nloc"], nhalo=res["nhalo"], iterations=res["stats"]["iterations"],
            solve_ms=res["stats"]["solve_ms"], first_solve_ms=res["first_solve_ms"], ms_per_iter=res["stats"]["solve_ms"] / max(res["stats"]["iterations"], 1),
            distributed=res["distributed"], device_mem_gb=res["device_mem_gb"], local_nodes=res["local_nodes"], local_tets=res["local_tets"],
            mesh_nodes=mesh.nn, mesh_tets=mesh.nt, true_rel_residual=res.get("true_rel_residual"),
            recurrence_rel_residual=res["stats"].get("recurrence_rel_residual"), **res["timings"])
if not distributed:
    line.update(rel_err_vs_single=res["rel_err_vs_single"], single_gpu_ms=res["single_gpu_ms"], single_gpu_iterations=res["single_gpu_iterations"],
                single_gpu_coarse_solve_ms=res["single_gpu_auto_solve_ms"], single_gpu_coarse_iterations=res["single_gpu_auto_iterations"])
sys.stdout.flush()
os.write(1, (json.dumps(line) + "\n").encode())     # one write: lines of different ranks do not interleave in the launcher's pipe
if os.environ.get("PTFEM_DUMP_X"):      # this rank's block of the solution, for the caller's own checks (tests compare with the oracle)
    np.save(os.environ["PTFEM_DUMP_X"].format(rank=rank), res["x_local"])
assert distributed or res["rel_err_vs_single"] < 1e-6, res["rel_err_vs_single"]
dist.destroy_process_group()
