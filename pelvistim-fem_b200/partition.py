"""Row partitioning of one assembled system for the multi-GPU solve (BASELINE.json config #5; not in
the reference, whose ElmerSolver run is serial: ``step03_ankle_layers/run_layered_sweep.py:1099``).

Rank *r* owns the contiguous rows ``[bounds[r], bounds[r+1])``.  Columns of its block are renumbered
``[0, nloc)`` = owned (global - row0) and ``[nloc, nloc+nhalo)`` = halo, the halo being the sorted
global ids of the non-owned columns (so it is automatically grouped by owner rank).  Because the
stiffness pattern is symmetric, the rows a neighbour needs from this rank are exactly this rank's
rows that have a column owned by that neighbour, in ascending order — no communication is needed to
build the send lists.

``cg_single_reduction`` is the host (numpy) statement of the iteration ``ptfem_dist_solve`` runs on
the GPUs (Chronopoulos-Gear: one halo exchange + one all-reduce of three scalars per iteration); the
CPU tests drive it over ``torch.distributed`` (gloo) to check the partitioning logic.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


def row_bounds(n, nranks):
    """Contiguous, balanced row ownership: int64 [nranks+1]."""
    return np.array([(n * r) // nranks for r in range(nranks + 1)], dtype=np.int64)


@dataclass
class LocalBlock:
    rank: int
    nranks: int
    row0: int
    nloc: int
    nhalo: int
    rowptr: np.ndarray        # int32 [nloc+1]
    col: np.ndarray           # int32 [nnz_loc], local numbering
    val: np.ndarray           # float64 [nnz_loc]
    b: np.ndarray             # float64 [nloc]
    halo_global: np.ndarray   # int64 [nhalo] global ids of the halo slots
    nbr_rank: np.ndarray      # int32 [nnbr]
    send_ptr: np.ndarray      # int32 [nnbr+1]
    send_idx: np.ndarray      # int32 [send_ptr[-1]] local rows to send, grouped by neighbour
    recv_ptr: np.ndarray      # int32 [nnbr+1] halo slots [nloc+recv_ptr[k], nloc+recv_ptr[k+1]) come from nbr k


def local_block(rowptr, col, val, b, rank, nranks, bounds=None) -> LocalBlock:
    n = rowptr.shape[0] - 1
    bounds = row_bounds(n, nranks) if bounds is None else np.asarray(bounds, dtype=np.int64)
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    nloc = r1 - r0
    k0, k1 = int(rowptr[r0]), int(rowptr[r1])
    lrowptr = (rowptr[r0:r1 + 1].astype(np.int64) - k0).astype(np.int32)
    c = col[k0:k1].astype(np.int64)
    owned = (c >= r0) & (c < r1)
    halo = np.unique(c[~owned])
    owner = np.searchsorted(bounds, halo, side="right") - 1
    lcol = np.where(owned, c - r0, nloc + np.searchsorted(halo, c)).astype(np.int32)
    nbr = np.unique(owner).astype(np.int32)
    recv_ptr = np.concatenate([[0], np.cumsum([(owner == q).sum() for q in nbr])]).astype(np.int32)
    # send lists: rows of this block with at least one column owned by q (pattern symmetry)
    rows = np.repeat(np.arange(nloc, dtype=np.int64), np.diff(lrowptr))
    col_owner = np.searchsorted(bounds, c, side="right") - 1
    send = []
    for q in nbr:
        send.append(np.unique(rows[col_owner == q]).astype(np.int32))
    send_ptr = np.concatenate([[0], np.cumsum([s.size for s in send])]).astype(np.int32)
    send_idx = np.concatenate(send).astype(np.int32) if send else np.zeros(0, np.int32)
    return LocalBlock(rank, nranks, r0, nloc, int(halo.size), lrowptr, lcol, np.ascontiguousarray(val[k0:k1], dtype=np.float64),
                      np.ascontiguousarray(b[r0:r1], dtype=np.float64), halo, nbr, send_ptr, send_idx, recv_ptr)


@dataclass
class LocalMesh:
    """One rank's part of a mesh for the distributed set-up: the owned nodes (global rows ``[row0, row0 + nloc)``, first),
    the ghost nodes of the elements that touch them (``halo_global``, ascending global id = grouped by owner) and, last, the
    remaining nodes of boundary triangles that touch any of those (``extra_global``: they only carry Dirichlet flags for
    the ghosts and end up as isolated rows).  Every tet with at least one owned node is here, so the owned ROWS of the
    stiffness matrix assembled on this mesh are complete, with columns already numbered [owned | halo]."""
    rank: int
    nranks: int
    row0: int
    nloc: int
    nn_global: int
    nodes: np.ndarray
    tets: np.ndarray
    region: np.ndarray
    tris: np.ndarray
    bcid: np.ndarray
    halo_global: np.ndarray
    extra_global: np.ndarray

    @property
    def nn(self):
        return self.nodes.shape[0]


def local_submesh(mesh, rank, nranks, bounds=None) -> LocalMesh:
    """Owner-computes sub-mesh of ``rank`` (assembly with ghost elements)."""
    nn = mesh.nodes.shape[0]
    bounds = row_bounds(nn, nranks) if bounds is None else np.asarray(bounds, dtype=np.int64)
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    nloc = r1 - r0
    t = mesh.tets
    sel = ((t >= r0) & (t < r1)).any(axis=1)
    lt = t[sel]
    ghosts = np.unique(lt[(lt < r0) | (lt >= r1)]).astype(np.int64)
    inloc = np.zeros(nn, dtype=bool)
    inloc[r0:r1] = True
    inloc[ghosts] = True
    tsel = inloc[mesh.tris].any(axis=1) if mesh.tris.shape[0] else np.zeros(0, dtype=bool)
    ltri = mesh.tris[tsel]
    extra = np.unique(ltri[~inloc[ltri]]).astype(np.int64) if ltri.size else np.zeros(0, dtype=np.int64)
    g2l = np.full(nn, -1, dtype=np.int32)
    g2l[r0:r1] = np.arange(nloc, dtype=np.int32)
    g2l[ghosts] = nloc + np.arange(ghosts.size, dtype=np.int32)
    g2l[extra] = nloc + ghosts.size + np.arange(extra.size, dtype=np.int32)
    nodes = np.ascontiguousarray(np.concatenate([mesh.nodes[r0:r1], mesh.nodes[ghosts], mesh.nodes[extra]], axis=0))
    return LocalMesh(rank, nranks, r0, nloc, nn, nodes, np.ascontiguousarray(g2l[lt]), np.ascontiguousarray(mesh.region[sel]),
                     np.ascontiguousarray(g2l[ltri]).reshape(-1, 3), np.ascontiguousarray(mesh.bcid[tsel]), ghosts, extra)


def block_from_local(lm: LocalMesh, rowptr, col, val, b, bounds=None) -> LocalBlock:
    """The rank's block of the row-partitioned system out of the matrix assembled on its :class:`LocalMesh`: rows
    ``[0, nloc)``, whose columns are already local ([0, nloc) owned, then the halo in ascending global order)."""
    bounds = row_bounds(lm.nn_global, lm.nranks) if bounds is None else np.asarray(bounds, dtype=np.int64)
    nloc, nh = lm.nloc, int(lm.halo_global.size)
    k1 = int(rowptr[nloc])
    lrowptr = np.ascontiguousarray(rowptr[:nloc + 1], dtype=np.int32)
    c = col[:k1].astype(np.int64)
    if c.size and int(c.max()) >= nloc + nh:
        raise ValueError("an owned row refers to a node outside [owned | halo]: the sub-mesh misses an element")
    owner = np.searchsorted(bounds, lm.halo_global, side="right") - 1
    nbr = np.unique(owner).astype(np.int32)
    recv_ptr = np.concatenate([[0], np.cumsum([(owner == q).sum() for q in nbr])]).astype(np.int32)
    rows = np.repeat(np.arange(nloc, dtype=np.int64), np.diff(lrowptr))
    col_owner = np.full(c.shape, lm.rank, dtype=np.int64)
    ish = c >= nloc
    col_owner[ish] = owner[c[ish] - nloc]
    send = [np.unique(rows[col_owner == q]).astype(np.int32) for q in nbr]
    send_ptr = np.concatenate([[0], np.cumsum([s_.size for s_ in send])]).astype(np.int32)
    send_idx = np.concatenate(send).astype(np.int32) if send else np.zeros(0, np.int32)
    return LocalBlock(lm.rank, lm.nranks, lm.row0, nloc, nh, lrowptr, np.ascontiguousarray(col[:k1], dtype=np.int32),
                      np.ascontiguousarray(val[:k1], dtype=np.float64), np.ascontiguousarray(b[:nloc], dtype=np.float64),
                      lm.halo_global, nbr, send_ptr, send_idx, recv_ptr)


class SerialComm:
    """All ranks in one process (tests): exchange/allreduce over a list of blocks."""

    def __init__(self, blocks):
        self.blocks = blocks

    def exchange_all(self, us):
        """us[r]: owned part of rank r's vector -> list of halo arrays."""
        out = []
        for blk in self.blocks:
            h = np.empty(blk.nhalo)
            for k, q in enumerate(blk.nbr_rank):
                peer = self.blocks[q]
                kk = int(np.nonzero(peer.nbr_rank == blk.rank)[0][0])
                h[blk.recv_ptr[k]:blk.recv_ptr[k + 1]] = us[q][peer.send_idx[peer.send_ptr[kk]:peer.send_ptr[kk + 1]]]
            out.append(h)
        return out


def check_consistency(blocks):
    """Send list of p towards q must be exactly q's halo slots owned by p, in order."""
    for blk in blocks:
        for k, q in enumerate(blk.nbr_rank):
            peer = blocks[q]
            kk = np.nonzero(peer.nbr_rank == blk.rank)[0]
            if kk.size != 1:
                return False
            kk = int(kk[0])
            want = blk.halo_global[blk.recv_ptr[k]:blk.recv_ptr[k + 1]]
            got = peer.row0 + peer.send_idx[peer.send_ptr[kk]:peer.send_ptr[kk + 1]].astype(np.int64)
            if not np.array_equal(want, got):
                return False
    return True


def halo_sources(blk: LocalBlock, bounds=None, n=None):
    """Owner-local row index mirrored by every halo slot (for the peer-memory pull): global id - owner's row0."""
    if bounds is None:
        bounds = row_bounds(n, blk.nranks)
    owner = np.searchsorted(np.asarray(bounds), blk.halo_global, side="right") - 1
    return (blk.halo_global - np.asarray(bounds)[owner]).astype(np.int32)


def local_spmv(blk: LocalBlock, u_full):
    """y_loc = A_loc [u_owned ; u_halo]."""
    import scipy.sparse as sp
    A = sp.csr_matrix((blk.val, blk.col, blk.rowptr), shape=(blk.nloc, blk.nloc + blk.nhalo))
    return A @ u_full


def cg_single_reduction(blk: LocalBlock, exchange, allreduce, rtol=1e-10, maxit=100000, precond=None):
    """Chronopoulos-Gear PCG on one rank's block.  ``exchange(u_owned) -> u_halo`` and ``allreduce(vec3) -> vec3`` are
    the only communication of the iteration itself; ``precond(r_owned, dinv) -> u_owned`` replaces the Jacobi step
    ``u = D^-1 r`` (it may communicate: the coarse-grid preconditioner sums its finest grid vector over the ranks, as
    ``csrc/dist.cu`` does).  Returns (x_owned, iterations, rel)."""
    nloc = blk.nloc
    diag = np.ones(nloc)
    rows = np.repeat(np.arange(nloc), np.diff(blk.rowptr))
    on_diag = blk.col == rows
    diag[rows[on_diag]] = blk.val[on_diag]
    dinv = 1.0 / diag
    if precond is None:
        precond = lambda r_, dinv_: r_ * dinv_
    x = np.zeros(nloc)
    r = blk.b.copy()
    u = precond(r, dinv)
    bn2 = allreduce(np.array([blk.b @ blk.b, 0.0, 0.0]))[0]
    w = local_spmv(blk, np.concatenate([u, exchange(u)]))
    g, d, rr = allreduce(np.array([r @ u, w @ u, r @ r]))
    alpha, beta, g_old = (g / d if d > 0 else 0.0), 0.0, g
    p = np.zeros(nloc)
    s = np.zeros(nloc)
    it = 0
    while it < maxit and rr > rtol * rtol * bn2:
        p = u + beta * p
        s = w + beta * s
        x += alpha * p
        r -= alpha * s
        u = precond(r, dinv)
        w = local_spmv(blk, np.concatenate([u, exchange(u)]))
        g, d, rr = allreduce(np.array([r @ u, w @ u, r @ r]))
        beta = g / g_old if g_old > 0 else 0.0
        den = d - beta * g / alpha if alpha != 0 else 0.0
        alpha = g / den if den > 0 else 0.0
        g_old = g
        it += 1
    return x, it, float(np.sqrt(rr / bn2)) if bn2 > 0 else 0.0
