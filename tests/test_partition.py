"""Row partitioning + single-reduction CG (host logic of the multi-GPU solve) on CPU: serial emulation of
all ranks, and a real 2-process run over torch.distributed (gloo)."""
import os

import numpy as np
import pytest

import pelvistim_fem_b200  # noqa: F401
from conftest import SIGMA5
from oracle import fem_oracle as fo
from pelvistim_fem_b200 import meshgen, partition


def _system():
    m = meshgen.synth_slab("XS")
    ref = fo.solve_case(m, SIGMA5, [(102, 0.0)], [(101, 15.975)], recover=None)
    K = ref["K"].tocsr()
    K.sort_indices()
    return m, ref, K.indptr.astype(np.int32), K.indices.astype(np.int32), K.data.copy(), ref["b"]


@pytest.mark.parametrize("nranks", [1, 2, 3, 8])
def test_blocks_are_consistent_and_spmv_matches(nranks):
    m, ref, rowptr, col, val, b = _system()
    # the oracle's eliminated matrix drops structural zeros: keep its own pattern (still symmetric)
    blocks = [partition.local_block(rowptr, col, val, b, r, nranks) for r in range(nranks)]
    assert partition.check_consistency(blocks)
    assert sum(bk.nloc for bk in blocks) == m.nn
    x = np.random.default_rng(0).standard_normal(m.nn)
    comm = partition.SerialComm(blocks)
    halos = comm.exchange_all([x[bk.row0:bk.row0 + bk.nloc] for bk in blocks])
    y = np.concatenate([partition.local_spmv(bk, np.concatenate([x[bk.row0:bk.row0 + bk.nloc], h])) for bk, h in zip(blocks, halos)])
    assert np.abs(y - ref["K"] @ x).max() < 1e-12 * np.abs(y).max()
    for bk in blocks:
        assert bk.col.max() < bk.nloc + bk.nhalo and bk.rowptr[-1] == bk.col.shape[0]
        assert np.all(np.diff(bk.halo_global) > 0)


def test_single_reduction_cg_one_rank_matches_direct():
    m, ref, rowptr, col, val, b = _system()
    blk = partition.local_block(rowptr, col, val, b, 0, 1)
    x, it, rel = partition.cg_single_reduction(blk, lambda u: np.zeros(0), lambda v: v, rtol=1e-12)
    assert rel <= 1e-12 and np.abs(x - ref["phi"]).max() < 1e-8 * np.abs(ref["phi"]).max()


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m, ref, rowptr, col, val, b = _system()
    blk = partition.local_block(rowptr, col, val, b, rank, world)

    def exchange(u):
        h = np.empty(blk.nhalo)
        reqs, bufs = [], []
        for k, qn in enumerate(blk.nbr_rank):
            s = torch.from_numpy(np.ascontiguousarray(u[blk.send_idx[blk.send_ptr[k]:blk.send_ptr[k + 1]]]))
            r = torch.empty(int(blk.recv_ptr[k + 1] - blk.recv_ptr[k]), dtype=torch.float64)
            reqs.append(dist.isend(s, int(qn)))
            reqs.append(dist.irecv(r, int(qn)))
            bufs.append((k, r, s))
        for rq in reqs:
            rq.wait()
        for k, r, _ in bufs:
            h[blk.recv_ptr[k]:blk.recv_ptr[k + 1]] = r.numpy()
        return h

    def allreduce(v):
        t = torch.from_numpy(v.copy())
        dist.all_reduce(t)
        return t.numpy()
    x, it, rel = partition.cg_single_reduction(blk, exchange, allreduce, rtol=1e-11)
    err = np.abs(x - ref["phi"][blk.row0:blk.row0 + blk.nloc]).max() / np.abs(ref["phi"]).max()
    q.put((rank, it, rel, err))
    dist.destroy_process_group()


def test_two_process_gloo_solve():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    its = {r[1] for r in res}
    assert len(its) == 1                                    # both ranks agree on the iteration count
    for rank, it, rel, err in res:
        assert rel <= 1e-11 and err < 1e-7, (rank, it, rel, err)


# -- sweep-level sharding: independent points over worker processes, rows gathered in sweep order ------------------
def _point_fn(pt):
    from pelvistim_fem_b200 import sweep
    return dict(point=pt, square=pt * pt, rank=sweep.worker_rank())


def _failing_fn(pt):
    if pt == 3:
        raise ValueError("boom")
    return pt


def test_sweep_map_points_two_workers():
    from pelvistim_fem_b200 import sweep
    out = sweep.map_points(_point_fn, list(range(7)), gpus=2)
    assert [r["point"] for r in out] == list(range(7)) and [r["square"] for r in out] == [k * k for k in range(7)]
    assert [r["rank"] for r in out] == [k % 2 for k in range(7)]            # point i -> GPU i mod G
    with pytest.raises(RuntimeError, match="boom"):
        sweep.map_points(_failing_fn, [1, 2, 3, 4], gpus=2)
