"""CPU ORACLE (test infrastructure only — never imported by the product path).

numpy/scipy restatement of the coarse-grid PCG preconditioner of ``pelvistim-fem_b200/csrc/coarse.cu``:

    M^-1 r = D^-1 r + w sum_l Z_l B_l Z_l^T r        (w = 2 / (levels + 1) unless ``PTFEM_COARSE_WEIGHT`` is set)

``Z_l``: trilinear interpolation from nested regular grids over the mesh bounding box to the mesh nodes (zero
rows for Dirichlet nodes), ``B`` of the coarsest grid = exact inverse of the Galerkin matrix ``Z^T K Z``, ``B_l`` of
the finer grids = inverse diagonal of theirs.  The reference has no counterpart (it solves directly with UMFPACK,
``step01_box/case.sif:41-42``): a preconditioner only changes how fast PCG converges, not the answer, so there is
nothing of the reference to pin here.  What the tests use it for: the solution equals the direct solve, and the
ITERATION COUNT of the CUDA solver equals this restatement's to within round-off — if the device built a wrong
Galerkin operator or a wrong interpolation, its count would be higher.

Grid choice, level count and the snapping of weights on grid planes follow coarse.cu (``choose_grid``,
``coarse_prepare``, ``coarse_locate``).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def choose_grid(lo, hi, target_nodes):
    """Cells per axis of the coarsest grid: about ``target_nodes`` nodes, near-cubic cells, >= 2 cells per axis."""
    ext = np.where(hi - lo > 0.0, hi - lo, 1.0)
    vol = float(np.prod(ext))
    h = (vol / target_nodes) ** (1.0 / 3.0)
    for _ in range(8):
        nodes = float(np.prod(np.maximum(2.0, np.round(ext / h)) + 1.0))
        h *= (nodes / target_nodes) ** (1.0 / 3.0)
    return np.maximum(2.0, np.round(ext / h)).astype(int)


def level_count(nn, base_cells, extra_levels):
    """Levels used: requested finer levels are dropped until >= 8 (requested) / 16 (automatic) nodes per finest cell."""
    cells = float(np.prod(base_cells))
    min_rows = 16.0 if extra_levels < 0 else 8.0
    nlev = 6 if extra_levels < 0 else min(extra_levels, 5) + 1
    while nlev > 1 and cells * 8.0 ** (nlev - 1) * min_rows > nn:
        nlev -= 1
    return nlev


def interpolation(nodes, free, lo, hi, n):
    """Z [nn, (n0+1)(n1+1)(n2+1)] for a grid of n cells per axis; rows of non-free nodes are zero."""
    ext = np.where(hi - lo > 0.0, hi - lo, 1.0)
    u = (nodes - lo) * (n / (ext * (1.0 + 1e-12)))
    c = np.clip(np.floor(u).astype(np.int64), 0, n - 1)
    t = u - c
    t = np.where(t < 1e-9, 0.0, np.where(t > 1.0 - 1e-9, 1.0, t))
    nn = nodes.shape[0]
    rows, cols, vals = [], [], []
    for a in range(8):
        ax, ay, az = a & 1, (a >> 1) & 1, a >> 2
        w = (t[:, 0] if ax else 1 - t[:, 0]) * (t[:, 1] if ay else 1 - t[:, 1]) * (t[:, 2] if az else 1 - t[:, 2])
        node = (c[:, 0] + ax) + (n[0] + 1) * ((c[:, 1] + ay) + (n[1] + 1) * (c[:, 2] + az))
        rows.append(np.arange(nn)); cols.append(node); vals.append(w * free)
    k = int(np.prod(n + 1))
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nn, k))


class CoarsePreconditioner:
    """M^-1 for the eliminated matrix K (Dirichlet rows = identity rows) of a mesh with nodes ``nodes``."""

    def __init__(self, K, nodes, is_dirichlet, coarse_nodes=300, extra_levels=-1, level_weight=None):
        lo, hi = nodes.min(axis=0), nodes.max(axis=0)
        base = choose_grid(lo, hi, float(coarse_nodes if coarse_nodes > 0 else 300))
        self.nlev = level_count(nodes.shape[0], base, extra_levels)
        free = (~np.asarray(is_dirichlet, dtype=bool)).astype(np.float64)
        # the additive levels overlap in what they correct: each is weighted by 2 / (levels + 1) against the Jacobi term
        # (coarse.cu: level_weight(); measured optimum on the 3-level bench mesh, neutral for one level)
        if level_weight is None:
            level_weight = 2.0 / (self.nlev + 1)
        self.level_weight = level_weight
        self.dinv = 1.0 / K.diagonal()
        self.Z, self.B = [], []
        for l in range(self.nlev):                      # level 0 = finest, last = coarsest (exact)
            n = base * (1 << (self.nlev - 1 - l))
            Z = interpolation(nodes, free, lo, hi, n)
            E = (Z.T @ K @ Z).tocsc()
            d = E.diagonal()
            if l == self.nlev - 1:
                Ed = E.toarray()
                empty = ~(d > 0.0)
                Ed[empty, empty] = 1.0
                self.B.append(level_weight * np.linalg.inv(Ed))
            else:
                self.B.append(level_weight * np.where(d > 0.0, 1.0 / np.where(d > 0.0, d, 1.0), 0.0))
            self.Z.append(Z)
        self.coarse_unknowns = self.Z[-1].shape[1]

    def apply(self, r):
        z = self.dinv * r
        for Z, B in zip(self.Z, self.B):
            rc = Z.T @ r
            z = z + Z @ (B @ rc if B.ndim == 2 else B * rc)
        return z


def pcg(K, b, apply_minv, rtol=1e-10, maxit=100000):
    """Preconditioned CG with the recurrences of the CUDA solver; returns (x, iterations)."""
    x = np.zeros_like(b)
    r = b.copy()
    z = apply_minv(r)
    p = z.copy()
    rz = r @ z
    bn = np.linalg.norm(b)
    it = 0
    while it < maxit and np.linalg.norm(r) > rtol * bn:
        q = K @ p
        alpha = rz / (p @ q)
        x += alpha * p
        r -= alpha * q
        z = apply_minv(r)
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
        it += 1
    return x, it
