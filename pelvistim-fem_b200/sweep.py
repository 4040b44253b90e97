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


class PipelinePool:
    """``pipelines`` host threads of THIS process, each owning a ``Context`` (its own streams) on the same GPU and a private
    ``state`` dict, kept alive between calls (the device allocator caches blocks per host thread).  Sweep points are
    independent (``run_layered_sweep.py:1061-1062``), so while one pipeline is in the latency-bound part of its point (mesh
    upload, pattern build, Python between calls) the other keeps the memory system busy with its solve: the GPU sees the
    union of both streams."""

    def __init__(self, device=0, pipelines=2):
        import queue
        import threading
        self.pipelines = max(1, int(pipelines))
        self.device = device
        self._jobs = [queue.Queue() for _ in range(self.pipelines)]
        self._done = queue.Queue()
        self._threads = [threading.Thread(target=self._run, args=(k,), name=f"ptfem-pipeline-{k}", daemon=True)
                         for k in range(self.pipelines)]
        for t in self._threads:
            t.start()

    def _run(self, k):
        from .engine import Context
        ctx, state = None, {"pipeline": k}
        while True:
            job = self._jobs[k].get()
            if job is None:
                break
            fn, items, finish, out = job
            err = None
            try:
                if ctx is None:
                    ctx = Context(self.device)
                for i, pt in items:
                    out[i] = fn(ctx, state, pt)
                if finish is not None:
                    finish(ctx, state)
                ctx.sync()
            except BaseException as e:  # noqa: BLE001 - re-raised in the caller's thread
                err = e
            self._done.put(err)
        if ctx is not None:
            ctx.close()
        self._done.put(None)

    def map(self, fn, points, finish=None):
        """``[fn(ctx, state, pt) for pt in points]``, point i on pipeline i mod P; ``finish(ctx, state)`` runs in every
        pipeline's thread after its last point (drain outstanding copies)."""
        points = list(points)
        out = [None] * len(points)
        for k in range(self.pipelines):
            self._jobs[k].put((fn, [(i, points[i]) for i in range(k, len(points), self.pipelines)], finish, out))
        errs = [self._done.get() for _ in range(self.pipelines)]
        for e in errs:
            if e is not None:
                raise e
        return out

    def close(self):
        for q in self._jobs:
            q.put(None)
        for _ in self._threads:
            self._done.get()
        for t in self._threads:
            t.join()
        self._threads = []


def map_points_pipelined(fn, points, device=0, pipelines=2, finish=None):
    """One-shot form of ``PipelinePool.map``."""
    pool = PipelinePool(device, pipelines)
    try:
        return pool.map(fn, points, finish)
    finally:
        pool.close()
