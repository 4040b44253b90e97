"""Launcher-side helper for the row-partitioned multi-GPU solve: every rank assembles the case on its own
GPU (mesh and pattern are replicated, as in a sweep), keeps the rows it owns, and the ranks then iterate
together over NCCL (``ptfem_dist_solve``).  Needs ``torch.distributed`` to be initialised (any backend)
for the one-off broadcast of the NCCL id."""
from __future__ import annotations

import time

import numpy as np

from . import engine, partition


def _allreduce(arr, op="sum"):
    """Element-wise reduction of a float64 numpy array over the ranks of the default process group (NCCL: through the
    GPU; gloo: on the host); every rank gets the same result."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return arr
    t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64))
    on_gpu = dist.get_backend() == "nccl"
    if on_gpu:
        t = t.cuda()
    dist.all_reduce(t, op={"sum": dist.ReduceOp.SUM, "min": dist.ReduceOp.MIN, "max": dist.ReduceOp.MAX}[op])
    return t.cpu().numpy() if on_gpu else t.numpy()


def partitioned_solve(ctx, mesh, sigma_by_body, dirichlet, neumann, rank, world, check=True, transport="p2p", coarse=True,
                      force_p2p=False, distributed=False, true_residual=True, **opts):
    """Returns dict(x_local, row0, stats, timings, nloc, nhalo, rel_err_vs_single).

    ``coarse``: precondition with Jacobi + geometric coarse grids (the finest grid vector is summed over the ranks once per
    iteration) when the mesh is large enough for the single-GPU solver to choose them too (>= 100 k nodes); Jacobi otherwise.
    ``force_p2p``: use the peer-memory kernels even with one rank (tests).
    ``true_residual``: recompute ||b - A x|| / ||b|| of the returned solution on the host from all ranks' rows
    (``true_rel_residual`` in the result and in ``stats``; the device loop converges on the recurrence residual).

    ``distributed=False``: every rank assembles a replica of the whole mesh on its GPU and cuts its row block out of it (the
    mesh must fit one GPU; with ``check`` the single-GPU solve of the replica is timed beside the partitioned one).
    ``distributed=True``: no GPU ever holds the whole mesh - each rank uploads its owned nodes and the elements touching
    them (``partition.local_submesh``), assembles that, and the coarse-grid operators are built from per-rank Galerkin
    sums added up over the ranks (owner-computes with ghost elements)."""
    import torch.distributed as dist
    lm = None
    if distributed:
        check = False
        lm = partition.local_submesh(mesh, rank, world)
        dm = ctx.mesh(lm.nodes, lm.tets, lm.region, lm.tris, lm.bcid)
        nn_global = lm.nn_global
        # bounding box of the WHOLE mesh (coarse grids are laid over it); also marks the local mesh as a part of a larger one
        own = lm.nodes[:lm.nloc]
        bb_lo, bb_hi = _allreduce(own.min(axis=0), "min"), _allreduce(own.max(axis=0), "max")
        dm.set_bbox(bb_lo, bb_hi)
    else:
        dm = ctx.mesh(mesh.nodes, mesh.tets, mesh.region, mesh.tris, mesh.bcid)
        nn_global = mesh.nn
    dm.assemble(sigma_by_body).bc_reset(1)
    for bid, g in neumann:
        dm.neumann(bid, g)
    for bid, v in dirichlet:
        dm.dirichlet(bid, v)
    rowptr, col = dm.get_pattern()
    val = dm.get_values(0, True)
    b = dm.get_rhs(0)
    phi_single = None
    if check:
        # same algorithm on one GPU (Jacobi-PCG) for the strong-scaling ratio, then the one-GPU default (coarse grids)
        # (each twice, the second reported: the first captures the CUDA graph of the iteration, like the partitioned solve below)
        dm.solve(to_host=False, rtol=opts.get("rtol", 1e-10), precond=engine.PRECOND_JACOBI)
        phi_single = dm.solve(rtol=opts.get("rtol", 1e-10), precond=engine.PRECOND_JACOBI)[0]
        single_stats = dm.last_stats
        dm.solve(to_host=False, rtol=opts.get("rtol", 1e-10), precond=engine.PRECOND_AUTO)
        setup_ms = dm.last_stats["setup_ms"]
        dm.solve(to_host=False, rtol=opts.get("rtol", 1e-10), precond=engine.PRECOND_AUTO)
        single_auto_stats = dict(dm.last_stats, setup_ms=setup_ms)
    if distributed:
        blk = partition.block_from_local(lm, rowptr, col, val, b)
    else:
        blk = partition.local_block(rowptr, col, val, b, rank, world)
    del rowptr, col, val, b
    want_coarse = bool(coarse) and nn_global >= 100000
    state = dict(coarse=False, note=None)
    attach_row0 = 0 if distributed else blk.row0
    if distributed and want_coarse:
        # coarse-grid operators of the WHOLE matrix from per-rank sums: grids over the global bounding box, Galerkin sums of
        # the owned rows, one all-reduce of ~1 MB, the same inversion on every rank (every rank takes part in every collective)
        sums, err = None, None
        try:
            sums = dm.coarse_partial(lm.nloc, nn_global)
        except engine.PtfemError as e:
            err = str(e)
        sizes = [None] * world
        dist.all_gather_object(sizes, None if sums is None else int(sums.size))
        if all(z is not None for z in sizes) and len(set(sizes)) == 1:
            sums = _allreduce(sums, "sum")
            try:
                dm.coarse_finish(sums)
            except engine.PtfemError as e:
                err = str(e)
        else:
            err = err or "the ranks disagree on the coarse grids"
        errs = [None] * world
        dist.all_gather_object(errs, err)
        if any(errs):
            want_coarse = False
            state["note"] = f"coarse grids not built ({[e for e in errs if e][0]}); Jacobi"

    def make_system():
        """This rank's block; with the coarse grids attached when asked for and possible (every rank decides alike: the
        replicas are identical).  Must run before the peer-memory export."""
        sysm = engine.DistSystem(ctx, blk)
        state["coarse"] = False
        if want_coarse:
            try:
                sysm.coarse_attach(dm, attach_row0)
                state["coarse"] = True
            except engine.PtfemError as e:
                state["note"] = f"coarse grids not attached ({e}); Jacobi"
        return sysm

    used = transport if (world > 1 or force_p2p) else "single"
    ds = None
    if (world > 1 or force_p2p) and transport == "p2p":
        # peer memory needs CUDA IPC between the ranks' processes; every rank must agree on whether it works
        # (every rank takes part in every collective below whatever failed locally: a rank that raised before a gather
        #  would otherwise pair its next collective with the others' previous one)
        p2p_error, exported = None, None
        try:
            engine.dist_init(ctx, None, rank, world)
            ds = make_system()
            my_ranges = ds.coarse_ranges().tolist() if state["coarse"] else None
        except engine.PtfemError as e:
            p2p_error = str(e)
            my_ranges = None
        # sharded coarse exchange: every rank learns which slab of the coarse grids every other rank's rows reach
        all_ranges = [None] * world
        dist.all_gather_object(all_ranges, my_ranges if p2p_error is None else None)
        try:
            if p2p_error is None:
                if state["coarse"] and all(r is not None for r in all_ranges):
                    ds.set_coarse_ranges(np.array(all_ranges, dtype=np.int64))
                exported = ds.p2p_export()
        except engine.PtfemError as e:
            p2p_error = str(e)
        handles = [None] * world
        dist.all_gather_object(handles, exported)                 # None = that rank could not export
        if p2p_error is None:
            if all(h is not None for h in handles):
                try:
                    ds.p2p_connect(handles, partition.halo_sources(blk, n=nn_global))
                except engine.PtfemError as e:
                    p2p_error = str(e)
            else:
                p2p_error = "export failed on another rank"
        flags = [None] * world
        dist.all_gather_object(flags, p2p_error)
        if any(flags):
            if ds is not None:
                ds.close()
                ds = None
            engine.dist_finalize(ctx)
            used = "nccl"
            if rank == 0:
                print(f"distsolve: peer-memory transport unavailable ({[f for f in flags if f][0]}); using NCCL", flush=True)
        else:
            dist.barrier()
    if world > 1 and used == "nccl":
        ids = [engine.dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        engine.dist_init(ctx, ids[0], rank, world)
    if ds is None:
        ds = make_system()

    def timed_solves(ds):
        """first solve: captures the CUDA graph of the iteration (and warms the peer mappings); the second one is reported.
        Every rank reports success or failure of its own solves; the ranks then agree (a timed-out peer-memory wait is
        pushed to every rank's mailbox, so all of them leave the solve within one read-back of the scalars)."""
        err, x, first_ms, wall = None, None, None, None
        try:
            ds.solve(**opts)
            first_ms = ds.last_stats["solve_ms"]
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            x = ds.solve(**opts)
            wall = time.perf_counter() - t0
        except engine.PtfemError as e:
            if e.code == engine.ERR_NOCONV:
                raise
            err = str(e)
        if world > 1:
            flags = [None] * world
            dist.all_gather_object(flags, err)
            bad = [f for f in flags if f]
            if bad:
                err = err or f"on another rank: {bad[0]}"
        return err, x, first_ms, wall

    mem_gb = None
    try:   # device memory in use on this rank with the local mesh, its matrix and the block all resident (peak of the set-up)
        import torch
        free_b, total_b = torch.cuda.mem_get_info(ctx.device)
        mem_gb = (total_b - free_b) / 2 ** 30
    except Exception:  # noqa: BLE001 - reporting only
        pass
    err, x, first_ms, wall = timed_solves(ds)
    if err and used == "p2p" and world > 1:
        # a peer stalled past the bounded wait: the peer-memory connection is out of step - start over on the NCCL transport
        if rank == 0:
            print(f"distsolve: peer-memory solve failed ({err}); retrying over NCCL", flush=True)
        ds.close()
        engine.dist_finalize(ctx)
        used = "nccl"
        ids = [engine.dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        engine.dist_init(ctx, ids[0], rank, world)
        ds = make_system()
        err, x, first_ms, wall = timed_solves(ds)
    dm.close()
    if err:
        raise engine.PtfemError(-3, err)
    out = dict(x_local=x, row0=blk.row0, nloc=blk.nloc, nhalo=blk.nhalo, stats=ds.last_stats, timings=ds.timings, wall_s=wall,
               transport=used, coarse=state["coarse"], coarse_note=state["note"], first_solve_ms=first_ms, distributed=bool(distributed),
               device_mem_gb=mem_gb, local_nodes=(lm.nn if lm is not None else mesh.nn), local_tets=(lm.tets.shape[0] if lm is not None else mesh.nt))
    if true_residual:
        # ||b - A x|| / ||b|| of what came back, recomputed on the host from every rank's rows (the device solver reports the
        # recurrence residual of the Chronopoulos-Gear iteration, which can drift from the true one on ill-conditioned systems)
        xg = np.zeros(nn_global)
        xg[blk.row0:blk.row0 + blk.nloc] = x
        xg = _allreduce(xg, "sum")
        u = np.concatenate([xg[blk.row0:blk.row0 + blk.nloc], xg[np.asarray(blk.halo_global, dtype=np.int64)]])
        r = blk.b - partition.local_spmv(blk, u)
        sums = _allreduce(np.array([float(r @ r), float(blk.b @ blk.b)]), "sum")
        out["true_rel_residual"] = float(np.sqrt(sums[0] / max(sums[1], 1e-300)))
        out["stats"] = dict(out["stats"], true_rel_residual=out["true_rel_residual"], recurrence_rel_residual=out["stats"].get("rel_residual"))
    if check:
        ref = phi_single[blk.row0:blk.row0 + blk.nloc]
        out["rel_err_vs_single"] = float(np.abs(x - ref).max() / max(np.abs(phi_single).max(), 1e-300))
        out["single_gpu_ms"] = single_stats["solve_ms"]
        out["single_gpu_iterations"] = single_stats["iterations"]
        out["single_gpu_auto_ms"] = single_auto_stats["solve_ms"] + single_auto_stats["setup_ms"]
        out["single_gpu_auto_solve_ms"] = single_auto_stats["solve_ms"]
        out["single_gpu_auto_setup_ms"] = single_auto_stats["setup_ms"]
        out["single_gpu_auto_precond"] = single_auto_stats["precond"]
        out["single_gpu_auto_iterations"] = single_auto_stats["iterations"]
    if world > 1:
        dist.barrier()          # nobody unmaps peer memory while a neighbour may still read it
    ds.close()
    if world > 1:
        engine.dist_finalize(ctx)
    return out
