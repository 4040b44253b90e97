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


# -- coarse-grid preconditioner of the partitioned solve (csrc/dist.cu: dist_coarse_add) -----------------------------
# Restated with the oracle's interpolation matrices: every rank restricts its OWN rows to the finest grid, that grid
# vector is summed over the ranks, the grid hierarchy above it is applied replicated (grid-to-grid transfers, exact
# coarsest solve), and every rank interpolates back to its own rows.
def _coarse_setup(coarse_nodes=36, extra_levels=1):
    from oracle import coarse_oracle as cz
    m, ref, rowptr, col, val, b = _system()
    K = ref["K"].tocsr()
    is_dir = np.zeros(m.nn, dtype=bool)
    is_dir[np.unique(m.tris[m.bcid == 102])] = True
    M = cz.CoarsePreconditioner(K, m.nodes, is_dir, coarse_nodes=coarse_nodes, extra_levels=extra_levels)
    lo, hi = m.nodes.min(axis=0), m.nodes.max(axis=0)
    base = cz.choose_grid(lo, hi, float(coarse_nodes))
    # P_l: trilinear interpolation from grid l to the nodes of grid l-1 (cells halve from level to level)
    P = [None]
    for l in range(1, M.nlev):
        nf = base * (1 << (M.nlev - l))
        ext = np.where(hi - lo > 0.0, hi - lo, 1.0) * (1.0 + 1e-12)
        ax = [lo[d] + ext[d] * np.arange(nf[d] + 1) / nf[d] for d in range(3)]
        Z, Y, X = np.meshgrid(ax[2], ax[1], ax[0], indexing="ij")          # x fastest, as the grid nodes are numbered
        pts = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
        P.append(cz.interpolation(pts, np.ones(pts.shape[0]), lo, hi, nf // 2))
    return m, ref, (rowptr, col, val, b), K, M, P


def _hierarchy(M, P, rc0):
    """finest grid residual (summed over the ranks) -> finest grid correction carrying all levels (replicated part)."""
    rc = [rc0]
    for l in range(1, M.nlev):
        rc.append(P[l].T @ rc[l - 1])
    y = [(B @ r if B.ndim == 2 else B * r) for B, r in zip(M.B, rc)]
    yt = y[-1]
    for l in range(M.nlev - 2, -1, -1):
        yt = y[l] + P[l + 1] @ yt
    return yt


def test_nested_grids_interpolate_exactly():
    # what the CUDA code relies on when it touches the mesh only on the finest level: Z_l = Z_{l-1} P_l
    m, ref, sysm, K, M, P = _coarse_setup()
    assert M.nlev == 2
    for l in range(1, M.nlev):
        d = (M.Z[l] - M.Z[l - 1] @ P[l])
        assert abs(d).max() < 1e-10            # weights are O(1); the box is widened by 1e-12 and planes snap at 1e-9


@pytest.mark.parametrize("nranks", [1, 2, 3])
def test_partitioned_coarse_preconditioner_equals_global(nranks):
    m, ref, (rowptr, col, val, b), K, M, P = _coarse_setup()
    bounds = partition.row_bounds(m.nn, nranks)
    r = np.random.default_rng(3).standard_normal(m.nn)
    parts = [M.Z[0][bounds[q]:bounds[q + 1]].T @ r[bounds[q]:bounds[q + 1]] for q in range(nranks)]
    rc0 = np.sum(parts, axis=0)                                    # the one cross-rank sum of the iteration
    assert np.abs(rc0 - M.Z[0].T @ r).max() < 1e-12 * np.abs(rc0).max()
    yt = _hierarchy(M, P, rc0)
    z = np.concatenate([M.dinv[bounds[q]:bounds[q + 1]] * r[bounds[q]:bounds[q + 1]] + M.Z[0][bounds[q]:bounds[q + 1]] @ yt
                        for q in range(nranks)])
    zg = M.apply(r)
    assert np.abs(z - zg).max() < 1e-11 * np.abs(zg).max()


def _gloo_coarse_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from oracle import coarse_oracle as cz
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m, ref, (rowptr, col, val, b), K, M, P = _coarse_setup()      # replicated set-up, as on the GPUs
    blk = partition.local_block(rowptr, col, val, b, rank, world)
    Zloc = M.Z[0][blk.row0:blk.row0 + blk.nloc]

    def allreduce(v):
        t = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64).copy())
        dist.all_reduce(t)
        return t.numpy()

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

    def precond(r, dinv):
        return dinv * r + Zloc @ _hierarchy(M, P, allreduce(Zloc.T @ r))

    x, it, rel = partition.cg_single_reduction(blk, exchange, allreduce, rtol=1e-11, precond=precond)
    _, it_jacobi, _ = partition.cg_single_reduction(blk, exchange, allreduce, rtol=1e-11)
    _, it_serial = cz.pcg(K, ref["b"], M.apply, rtol=1e-11)
    err = np.abs(x - ref["phi"][blk.row0:blk.row0 + blk.nloc]).max() / np.abs(ref["phi"]).max()
    q.put((rank, it, it_jacobi, it_serial, rel, err))
    dist.destroy_process_group()


def test_two_process_gloo_solve_with_coarse_grids():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_coarse_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    for rank, it, it_jacobi, it_serial, rel, err in res:
        assert rel <= 1e-11 and err < 1e-8
        assert abs(it - it_serial) <= 2            # same preconditioner, same Krylov space: the partitioned count is the serial one
        assert it * 4 < it_jacobi * 3                # a 36-node coarse grid on 2254 nodes: 101 vs 157 iterations
    assert res[0][1] == res[1][1]


def test_local_submesh_gives_complete_owned_rows():
    # distributed set-up (owner-computes with ghost elements): the matrix assembled on a rank's sub-mesh has
    # complete owned rows with columns already numbered [owned | halo], and the blocks of all ranks fit together
    import scipy.sparse as sp
    from oracle import fem_oracle as fo
    from pelvistim_fem_b200 import meshgen
    m = meshgen.synth_slab("XS")
    sig = {1: 0.35, 2: 0.04, 3: 0.001, 4: 0.005, 5: 0.005}
    ref = fo.solve_case(m, sig, [(102, 0.0)], [(101, 15.975)], recover=None)
    K, b = ref["K"].tocsr(), ref["b"]

    def on_pattern(A, nodes, tets):
        """A's values on the FULL P1 pattern (explicit zeros where the elimination left none): what the device stores."""
        rp, cc = fo.csr_pattern(nodes.shape[0], tets)
        rows = np.repeat(np.arange(nodes.shape[0]), np.diff(rp))
        return rp, cc, np.asarray(A[rows, cc]).ravel()

    grp, gcc, gval = on_pattern(K, m.nodes, m.tets)
    for world in (2, 3):
        blocks = []
        for r in range(world):
            lm = partition.local_submesh(m, r, world)
            assert lm.nloc == partition.row_bounds(m.nn, world)[r + 1] - lm.row0 and lm.nn == lm.nloc + lm.halo_global.size + lm.extra_global.size
            Kl = fo.assemble_stiffness(lm.nodes, lm.tets, lm.region, sig).tocsr()
            isd, dv = fo.dirichlet_nodes(lm.tris, lm.bcid, [(102, 0.0)], lm.nn)
            bl = fo.neumann_rhs(lm.nodes, lm.tris, lm.bcid, [(101, 15.975)])
            Kl, bl = fo.apply_dirichlet_symmetric(Kl, bl, isd, dv)
            rp, cc, val = on_pattern(Kl.tocsr(), lm.nodes, lm.tets)
            blk = partition.block_from_local(lm, rp, cc, val, bl)
            want = partition.local_block(grp, gcc, gval, b, r, world)
            assert np.array_equal(blk.rowptr, want.rowptr) and np.array_equal(blk.halo_global, want.halo_global)
            # same rows (the sub-mesh numbers its halo after the owned nodes, so the order of the columns inside a row differs)
            A1 = sp.csr_matrix((blk.val, blk.col, blk.rowptr), shape=(blk.nloc, blk.nloc + blk.nhalo))
            A2 = sp.csr_matrix((want.val, want.col, want.rowptr), shape=(blk.nloc, blk.nloc + blk.nhalo))
            assert abs(A1 - A2).max() < 1e-15 * np.abs(want.val).max() and np.abs(blk.b - want.b).max() < 1e-18
            P1 = sp.csr_matrix((np.ones(blk.col.size), blk.col, blk.rowptr), shape=A1.shape)
            P2 = sp.csr_matrix((np.ones(want.col.size), want.col, want.rowptr), shape=A1.shape)
            assert (P1 != P2).nnz == 0                                  # identical pattern, explicit zeros included
            assert np.array_equal(blk.nbr_rank, want.nbr_rank) and np.array_equal(blk.send_idx, want.send_idx) and np.array_equal(blk.recv_ptr, want.recv_ptr)
            blocks.append(blk)
        assert partition.check_consistency(blocks)
