"""Elmer ASCII mesh directory reader / writer (the wire format between the
reference's ``ElmerGrid 14 2 mesh.msh -out elmer_mesh`` step and its
``ElmerSolver case.sif`` step).

File formats as parsed/written by the reference itself
(``step01_box/find_boundaries.py:16-40,87-90,104-108``,
``step01_box/setup_case.py:18-25,57-75``):

    mesh.header    nNodes nElems nBoundary / nTypes / <type> <count> ...
    mesh.nodes     id  -1  x y z
    mesh.elements  id  body  504  n1 n2 n3 n4
    mesh.boundary  id  bc  parent1 parent2  303  n1 n2 n3

Ids are 1-based in the files and 0-based in memory.
"""
from __future__ import annotations

import os
from pathlib import Path

import numpy as np

from .meshgen import TetMesh


def write_elmer_mesh(mesh_dir, mesh: TetMesh):
    d = Path(mesh_dir)
    d.mkdir(parents=True, exist_ok=True)
    nn, nt, nb = mesh.nn, mesh.nt, mesh.nb
    with open(d / "mesh.header", "w") as f:
        f.write(f"{nn} {nt} {nb}\n2\n504 {nt}\n303 {nb}\n")
    ids = np.arange(1, nn + 1)
    with open(d / "mesh.nodes", "w") as f:
        np.savetxt(f, np.column_stack([ids, -np.ones(nn), mesh.nodes]),
                   fmt=["%d", "%d", "%.17g", "%.17g", "%.17g"])
    with open(d / "mesh.elements", "w") as f:
        np.savetxt(f, np.column_stack([np.arange(1, nt + 1), mesh.region, np.full(nt, 504), mesh.tets + 1]),
                   fmt="%d")
    parent = mesh.tri_parent if mesh.tri_parent is not None else np.full(nb, -1, dtype=np.int32)
    with open(d / "mesh.boundary", "w") as f:
        np.savetxt(f, np.column_stack([np.arange(1, nb + 1), mesh.bcid, parent + 1, np.zeros(nb, dtype=np.int64),
                                       np.full(nb, 303), mesh.tris + 1]), fmt="%d")


def _read_table(path):
    """Read a whitespace table whose rows may have different lengths."""
    with open(path) as f:
        txt = f.read()
    rows = [ln.split() for ln in txt.splitlines() if ln.strip() and not ln.lstrip().startswith("!")]
    return rows


def read_elmer_mesh(mesh_dir) -> TetMesh:
    """Read an Elmer mesh directory holding linear tets (504) and boundary
    triangles (303).  Other element types raise ``ValueError`` (the reference
    only ever produces 504/303: ``find_boundaries.py:104-108``)."""
    d = Path(mesh_dir)
    for name in ("mesh.nodes", "mesh.elements", "mesh.boundary"):
        if not (d / name).exists():
            raise FileNotFoundError(f"{d / name} not found")
    # nodes: id tag x y z ("coords are always the last 3", setup_case.py:24)
    raw = np.loadtxt(d / "mesh.nodes", ndmin=2)
    nid = raw[:, 0].astype(np.int64)
    xyz = np.ascontiguousarray(raw[:, -3:], dtype=np.float64)
    nn = nid.shape[0]
    contiguous = np.array_equal(nid, np.arange(1, nn + 1))
    if not contiguous:
        lut = np.full(nid.max() + 1, -1, dtype=np.int64)
        lut[nid] = np.arange(nn)
    el = np.loadtxt(d / "mesh.elements", dtype=np.int64, ndmin=2)
    if el.shape[1] != 7 or not (el[:, 2] == 504).all():
        raise ValueError("only linear tetrahedra (Elmer type 504) are supported")
    region = el[:, 1].astype(np.int32)
    tets = el[:, 3:7]
    rows = _read_table(d / "mesh.boundary")
    tri_rows = [r for r in rows if len(r) >= 8 and r[4] == "303"]
    skipped = len(rows) - len(tri_rows)
    if tri_rows:
        b = np.array([r[:8] for r in tri_rows], dtype=np.int64)
    else:
        b = np.zeros((0, 8), dtype=np.int64)
    bcid = b[:, 1].astype(np.int32)
    parent = b[:, 2] - 1
    tris = b[:, 5:8]
    if contiguous:
        tets = tets - 1
        tris = tris - 1
    else:
        tets = lut[tets]
        tris = lut[tris]
    if nn and (tets.min() < 0 or tets.max() >= nn or (tris.size and (tris.min() < 0 or tris.max() >= nn))):
        raise ValueError("element refers to an unknown node id")
    return TetMesh(xyz, np.ascontiguousarray(tets, dtype=np.int32), region,
                   np.ascontiguousarray(tris, dtype=np.int32), bcid,
                   tri_parent=parent.astype(np.int32), meta=dict(kind="elmer", skipped_boundary=skipped,
                                                                 path=os.fspath(d)))
