"""Sweep-level data parallelism: the reference's sweep loops run their independent points one after
another (``run_layered_sweep.py:1061-1062``, ``run_pressure_sweep.py:709``); here point *i* goes to GPU
*i mod G*.  One worker process per GPU (each owns a ``Context``), no data-path collective; the parent
gathers the result rows in sweep order."""
from __future__ import annotations

import multiprocessing as mp
import os

_WORKER = {"rank": None, "ctx": None}


def worker_rank():
    return _WORKER["rank"]


def worker_context():
    """The calling process's GPU context (device = worker rank, or 0 in the parent)."""
    if _WORKER["ctx"] is None:
        from .engine import Context
        _WORKER["ctx"] = Context(_WORKER["rank"] or 0)
    return _WORKER["ctx"]


def assign(n_points, n_gpus):
    """Round-robin ownership: list (per GPU) of point indices."""
    return [list(range(g, n_points, n_gpus)) for g in range(n_gpus)]


def _worker(rank, fn, points, idx, out_q):
    _WORKER["rank"] = rank
    try:
        from .engine import bind_host_to_gpu
        bind_host_to_gpu(rank)                # host buffers of this worker live on its GPU's NUMA node
        for i in idx:
            out_q.put((i, fn(points[i]), None))
    except BaseException as e:  # noqa: BLE001 - reported to the parent, which raises
        out_q.put((-1, None, f"worker {rank}: {type(e).__name__}: {e}"))


def map_points(fn, points, gpus=1):
    """``[fn(pt) for pt in points]`` with the points sharded over ``gpus`` worker processes."""
    points = list(points)
    if gpus <= 1 or len(points) <= 1:
        return [fn(pt) for pt in points]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(g, fn, points, idx, q)) for g, idx in enumerate(assign(len(points), gpus)) if idx]
    for pr in procs:
        pr.start()
    out = [None] * len(points)
    got = 0
    err = None
    import queue
    while got < len(points) and err is None:
        try:
            i, res, e = q.get(timeout=2.0)
        except queue.Empty:
            # a worker killed by a CUDA abort or a segfault never reports: notice it instead of waiting forever
            dead = [pr for pr in procs if not pr.is_alive() and pr.exitcode not in (0, None)]
            if dead:
                err = f"sweep worker pid {dead[0].pid} died with exit code {dead[0].exitcode}"
            continue
        if e is not None:
            err = e
            break
        out[i] = res
        got += 1
    for pr in procs:
        if err is not None:
            pr.terminate()
        pr.join()
    if err is not None:
        raise RuntimeError(err)
    return out
