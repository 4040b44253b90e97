"""CPU ORACLE (test infrastructure only — never imported by the product path).

Float64 numpy/scipy restatement of the arithmetic the reference delegates to
``ElmerSolver`` (``Procedure = "StatCurrentSolve" "StatCurrentSolver"``,
``step01_box/case.sif:33-45``): P1 tetrahedral assembly of div(sigma grad phi) = 0,
``Current Density`` (Neumann) and ``Potential`` (Dirichlet) boundary conditions
(``case.sif:61-71``, ``step03_ankle_layers/run_layered_sweep.py:595-624``), a direct
sparse solve (``Linear System Direct Method = UMFPACK`` -> SuperLU here) and the
nodal ``volume current`` recovery (``Calculate Volume Current = True``,
``case.sif:39``).

PARITY STATUS: the algorithm lives in Elmer FEM (third-party, un-vendored,
version unpinned; no source under /root/reference) and cannot be run here.  The
oracle is pinned by the reference's analytic known-answer test
(``step01_box/test_step01_baseline.py:59-104``: phi = z/Lz, J = (0,0,-sigma/Lz), exact
for P1 on any tet mesh) and, to discretisation accuracy only, by the committed
step03/step04 ``summary.csv`` tables.  Node-level fields of the reference are
"parity unpinned" (no mesh / VTU of the reference exists on disk).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg
may import this module.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


# -- element geometry ----------------------------------------------------------
def tet_geometry(nodes, tets):
    """Volume and shape-function gradients of linear tets.

    N_i(x) = a_i + g_i . x ;  returns vol [nt] (signed, >0 for positive orientation)
    and grad [nt, 4, 3]."""
    p = nodes[tets]                                     # [nt,4,3]
    Jm = np.stack([p[:, 1] - p[:, 0], p[:, 2] - p[:, 0], p[:, 3] - p[:, 0]], axis=1)   # rows = edges
    det = np.linalg.det(Jm)
    inv = np.linalg.inv(Jm)                             # columns of inv = grad N1..N3
    g = np.empty((tets.shape[0], 4, 3))
    g[:, 1:, :] = inv.transpose(0, 2, 1)
    g[:, 0, :] = -g[:, 1:, :].sum(axis=1)
    return det / 6.0, g


def csr_pattern(nn, tets):
    """Canonical CSR sparsity pattern of the P1 stiffness matrix: rows in mesh node
    order, columns sorted ascending, diagonal present (node-node adjacency through
    shared tets).  Returns (rowptr int32 [nn+1], col int32 [nnz])."""
    i = np.repeat(tets, 4, axis=1).ravel()
    j = np.tile(tets, (1, 4)).ravel()
    key = np.unique(i.astype(np.int64) * nn + j.astype(np.int64))
    # nodes not referenced by any tet still get their diagonal entry
    diag = np.arange(nn, dtype=np.int64) * nn + np.arange(nn, dtype=np.int64)
    key = np.union1d(key, diag)
    rows = key // nn
    col = (key % nn).astype(np.int32)
    rowptr = np.zeros(nn + 1, dtype=np.int64)
    np.add.at(rowptr, rows + 1, 1)
    return np.cumsum(rowptr).astype(np.int32), col


def _sigma_per_tet(region, sigma_by_body):
    sig = np.empty(region.shape[0])
    seen = np.zeros(region.shape[0], dtype=bool)
    for body, s in sigma_by_body.items():
        m = region == body
        sig[m] = s
        seen |= m
    if not seen.all():
        raise ValueError(f"no conductivity for bodies {sorted(set(region[~seen].tolist()))}")
    return sig


def assemble_stiffness(nodes, tets, region, sigma_by_body):
    """K_ij = sum_e sigma_e V_e grad N_i . grad N_j  (CSR, canonical pattern)."""
    nn = nodes.shape[0]
    vol, g = tet_geometry(nodes, tets)
    sig = _sigma_per_tet(region, sigma_by_body)
    ke = (sig * vol)[:, None, None] * np.einsum("eik,ejk->eij", g, g)     # [nt,4,4]
    return _scatter(nn, tets, ke)


def assemble_mass(nodes, tets):
    """Consistent P1 mass matrix M_ij = sum_e V_e (1 + delta_ij) / 20."""
    nn = nodes.shape[0]
    vol, _ = tet_geometry(nodes, tets)
    me = vol[:, None, None] * ((np.ones((4, 4)) + np.eye(4)) / 20.0)[None]
    return _scatter(nn, tets, me)


def _scatter(nn, tets, ke):
    """Sum element matrices into the canonical pattern (structural zeros are kept:
    the pattern is the node adjacency, not the set of non-zero values)."""
    rowptr, col = csr_pattern(nn, tets)
    rows = np.repeat(np.arange(nn, dtype=np.int64), np.diff(rowptr))
    key = rows * nn + col
    i = np.repeat(tets, 4, axis=1).ravel().astype(np.int64)
    j = np.tile(tets, (1, 4)).ravel().astype(np.int64)
    pos = np.searchsorted(key, i * nn + j)
    val = np.bincount(pos, weights=ke.ravel(), minlength=col.shape[0])
    return sp.csr_matrix((val, col, rowptr), shape=(nn, nn))


# -- boundary conditions ------------------------------------------------------
def tri_areas(nodes, tris):
    p = nodes[tris]
    return 0.5 * np.linalg.norm(np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]), axis=1)


def neumann_rhs(nodes, tris, bcid, neumann):
    """b_i += g * A_tri / 3 for each node of each boundary triangle with a
    ``Current Density = g`` BC; positive g drives current INTO the domain
    (``bc_debug_report.txt:17-19``)."""
    b = np.zeros(nodes.shape[0])
    area = tri_areas(nodes, tris)
    for bid, gval in neumann:
        m = bcid == bid
        contrib = np.repeat(gval * area[m] / 3.0, 3)
        np.add.at(b, tris[m].ravel(), contrib)
    return b


def dirichlet_nodes(tris, bcid, dirichlet, nn):
    """Every node of a boundary element with the target id gets ``Potential = v``;
    later BCs override earlier ones.  Returns (is_dir bool[nn], value f64[nn])."""
    is_dir = np.zeros(nn, dtype=bool)
    val = np.zeros(nn)
    for bid, v in dirichlet:
        idx = np.unique(tris[bcid == bid].ravel())
        is_dir[idx] = True
        val[idx] = v
    return is_dir, val


def apply_dirichlet_symmetric(K, b, is_dir, val):
    """Symmetric elimination: b -= K[:,D] phi_D ; zero rows/cols of D ; diag = 1 ; b_D = phi_D.
    (Elmer zeroes the row only; same solution.)"""
    K = K.tocsr().copy()
    b = b - K @ (val * is_dir)
    keep = (~is_dir).astype(np.float64)
    Dk = sp.diags(keep)
    K2 = (Dk @ K @ Dk + sp.diags(is_dir.astype(np.float64))).tocsr()
    b2 = np.where(is_dir, val, b)
    K2.sort_indices()
    return K2, b2


def solve_direct(K, b):
    lu = spla.splu(K.tocsc(), permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0,
                   options=dict(SymmetricMode=True))
    return lu.solve(b)


# -- post-processing ----------------------------------------------------------
def element_fields(nodes, tets, region, sigma_by_body, phi):
    """Per tet: E_e = -sum_i phi_i grad N_i, J_e = sigma_e E_e."""
    vol, g = tet_geometry(nodes, tets)
    E = -np.einsum("ei,eik->ek", phi[tets], g)
    sig = _sigma_per_tet(region, sigma_by_body)
    return E, sig[:, None] * E, vol


def recover_nodal_current(nodes, tets, region, sigma_by_body, phi, method="l2"):
    """Nodal ``volume current`` J = -sigma grad phi.

    method "l2"      : Galerkin L2 projection  M J^k = int (-sigma d_k phi) N_i dV  (consistent mass)
           "lumped"  : same right-hand side, row-sum lumped mass (= volume-weighted mean of J_e)
           "average" : unweighted mean of J_e over the tets touching the node.
    Which of these Elmer's ``Calculate Volume Current`` implements cannot be read here (no Elmer source); the
    reference's own step03/step04 tables favour "lumped": the pad current integrated from the nodal field is
    5.58 / 5.30 / 5.20 mA (r = 5 / 10 / 15 mm) with it and 6.25 / 5.65 / 5.39 mA with "l2", against
    5.51 / 5.27 / 5.14 mA in ``step03_ankle_layers/results/summary.csv``."""
    nn = nodes.shape[0]
    _, Je, vol = element_fields(nodes, tets, region, sigma_by_body, phi)
    if method == "average":
        acc = np.zeros((nn, 3))
        cnt = np.zeros(nn)
        for a in range(4):
            np.add.at(acc, tets[:, a], Je)
            np.add.at(cnt, tets[:, a], 1.0)
        return acc / np.maximum(cnt, 1.0)[:, None]
    rhs = np.zeros((nn, 3))
    w = (vol / 4.0)[:, None] * Je
    for a in range(4):
        np.add.at(rhs, tets[:, a], w)
    if method == "lumped":
        ml = np.zeros(nn)
        for a in range(4):
            np.add.at(ml, tets[:, a], vol / 4.0)
        return rhs / np.where(ml > 0, ml, 1.0)[:, None]
    M = assemble_mass(nodes, tets)
    d = M.diagonal()
    if (d <= 0).any():                                 # unreferenced nodes
        M = M + sp.diags((d <= 0).astype(np.float64))
    lu = spla.splu(M.tocsc(), permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0,
                   options=dict(SymmetricMode=True))
    return np.stack([lu.solve(rhs[:, k]) for k in range(3)], axis=1)


def solve_case(mesh, sigma_by_body, dirichlet, neumann, recover="l2"):
    """Full restated solve step: returns dict(phi, J, K (with BCs), b, K_raw, b_neumann)."""
    nn = mesh.nodes.shape[0]
    K_raw = assemble_stiffness(mesh.nodes, mesh.tets, mesh.region, sigma_by_body)
    b_n = neumann_rhs(mesh.nodes, mesh.tris, mesh.bcid, neumann)
    is_dir, val = dirichlet_nodes(mesh.tris, mesh.bcid, dirichlet, nn)
    if not is_dir.any():
        raise ValueError("pure Neumann problem (no Potential BC) is singular")
    K, b = apply_dirichlet_symmetric(K_raw, b_n, is_dir, val)
    phi = solve_direct(K, b)
    J = recover_nodal_current(mesh.nodes, mesh.tets, mesh.region, sigma_by_body, phi, recover) if recover else None
    return dict(phi=phi, J=J, K=K, b=b, K_raw=K_raw, b_neumann=b_n, is_dir=is_dir, dir_val=val)


def jacobi_pcg(K, b, rtol=1e-12, maxit=100000, x0=None):
    """Reference Jacobi-preconditioned CG (same recurrences as the CUDA solver), numpy."""
    d = K.diagonal()
    x = np.zeros_like(b) if x0 is None else x0.copy()
    r = b - K @ x
    z = r / d
    p = z.copy()
    rz = r @ z
    bn = np.linalg.norm(b)
    it = 0
    while it < maxit:
        if np.linalg.norm(r) <= rtol * bn:
            break
        q = K @ p
        alpha = rz / (p @ q)
        x += alpha * p
        r -= alpha * q
        z = r / d
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
        it += 1
    return x, it
