"""Gmsh ``.msh`` reader (formats 2.2 and 4.1, ASCII) and 2.2 writer, and the ``ElmerGrid 14 2`` conversion
the reference runs between meshing and solving
(``subprocess ["ElmerGrid","14","2","mesh.msh","-out","elmer_mesh"]``: ``step03_ankle_layers/run_layered_sweep.py:1077``,
``step02_electrodes/run_sweep.py:315``, ``step04_pressure/run_pressure_sweep.py:696``,
``step01_box/test_step01_baseline.py:49``).  ``gmsh.write("mesh.msh")`` produces MSH 4.1 ASCII by default
(``run_layered_sweep.py:342-343``); ``gmsh -3 box.geo -o box.msh`` likewise.

Only what the reference's meshes contain is supported: linear tetrahedra (Gmsh type 4) carrying the physical
volume id (-> Elmer body) and 3-node triangles (type 2) carrying the physical surface id (-> Elmer boundary);
points and lines are dropped, as ElmerGrid does for a 3-D mesh.  Ids: kept as they are when the file names its
physical groups (``$PhysicalNames``, the step02-04 meshes -> boundaries 101/102/103), renumbered 1..K in
ascending order when it does not (``box.geo`` -> boundaries 1/2/3; evidence ``step01_box/case.sif:63,69``).
ElmerGrid itself is not available here, so this mapping is a restatement, not a verified copy.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

from .meshgen import TetMesh, find_tri_parents, orient_positive


def _sections(text):
    out, name, buf = {}, None, []
    for ln in text.splitlines():
        s = ln.strip()
        if s.startswith("$End"):
            out[name] = buf
            name, buf = None, []
        elif s.startswith("$") and name is None:
            name, buf = s[1:], []
        elif name is not None:
            buf.append(s)
    return out


def read_msh(path, renumber=None) -> TetMesh:
    sec = _sections(Path(path).read_text())
    if "MeshFormat" not in sec:
        raise ValueError(f"{path}: not a Gmsh .msh file")
    fmt = sec["MeshFormat"][0].split()
    version, binary = float(fmt[0]), int(fmt[1])
    if binary:
        raise ValueError("binary .msh files are not supported (the reference writes ASCII)")
    named = bool(sec.get("PhysicalNames")) and int(sec["PhysicalNames"][0]) > 0
    if version < 3.0:
        tags, xyz, tets, treg, tris, tbc = _read_v2(sec)
    elif version >= 4.0:
        tags, xyz, tets, treg, tris, tbc = _read_v4(sec)
    else:
        raise ValueError(f"unsupported MSH version {version}")
    if tets.shape[0] == 0:
        raise ValueError(f"{path}: no tetrahedra (Gmsh type 4) with a physical volume found")
    # compact numbering over the nodes the elements use, ascending original tag
    used = np.unique(np.concatenate([tets.ravel(), tris.ravel()]))
    order = np.argsort(tags, kind="stable")
    pos = np.searchsorted(tags[order], used)
    if not np.array_equal(tags[order][pos], used):
        raise ValueError("element refers to an unknown node tag")
    nodes = xyz[order][pos]
    lut = {int(t): i for i, t in enumerate(used.tolist())}
    remap = np.vectorize(lut.__getitem__, otypes=[np.int64])
    tets = remap(tets).astype(np.int32) if tets.size else tets.astype(np.int32)
    tris = remap(tris).astype(np.int32) if tris.size else np.zeros((0, 3), np.int32)
    if renumber is None:
        renumber = not named
    if renumber:
        treg = (np.searchsorted(np.unique(treg), treg) + 1).astype(np.int32)
        if tbc.size:
            tbc = (np.searchsorted(np.unique(tbc), tbc) + 1).astype(np.int32)
    orient_positive(nodes, tets)
    m = TetMesh(np.ascontiguousarray(nodes), np.ascontiguousarray(tets), treg.astype(np.int32), np.ascontiguousarray(tris),
                tbc.astype(np.int32), meta=dict(kind="gmsh", version=version, named=named, path=str(path)))
    m.tri_parent = find_tri_parents(m.tets, m.tris)
    return m


def _read_v2(sec):
    L = sec["Nodes"]
    n = int(L[0])
    a = np.array([ln.split() for ln in L[1:1 + n]], dtype=np.float64)
    tags, xyz = a[:, 0].astype(np.int64), a[:, 1:4]
    tets, treg, tris, tbc = [], [], [], []
    E = sec["Elements"]
    for ln in E[1:1 + int(E[0])]:
        f = ln.split()
        etype, ntags = int(f[1]), int(f[2])
        phys = int(f[3]) if ntags > 0 else 0
        nod = [int(v) for v in f[3 + ntags:]]
        if etype == 4 and phys > 0:
            tets.append(nod[:4]); treg.append(phys)
        elif etype == 2 and phys > 0:
            tris.append(nod[:3]); tbc.append(phys)
    return (tags, xyz, np.array(tets, dtype=np.int64).reshape(-1, 4), np.array(treg, dtype=np.int64),
            np.array(tris, dtype=np.int64).reshape(-1, 3), np.array(tbc, dtype=np.int64))


def _read_v4(sec):
    # entities: physical tags of surfaces (dim 2) and volumes (dim 3)
    phys = {2: {}, 3: {}}
    if "Entities" in sec:
        L = sec["Entities"]
        npnt, ncur, nsur, nvol = (int(v) for v in L[0].split())
        i = 1 + npnt + ncur
        for dim, cnt in ((2, nsur), (3, nvol)):
            for ln in L[i:i + cnt]:
                f = ln.split()
                nph = int(f[7])
                # (a negative physical tag only says the entity enters the group with reversed orientation)
                phys[dim][int(f[0])] = [abs(int(v)) for v in f[8:8 + nph]]
            i += cnt
    L = sec["Nodes"]
    nblocks, nn = int(L[0].split()[0]), int(L[0].split()[1])
    tags = np.empty(nn, dtype=np.int64)
    xyz = np.empty((nn, 3))
    i, k = 1, 0
    for _ in range(nblocks):
        cnt = int(L[i].split()[3])
        i += 1
        tags[k:k + cnt] = [int(v) for v in L[i:i + cnt]]
        i += cnt
        xyz[k:k + cnt] = [[float(v) for v in ln.split()[:3]] for ln in L[i:i + cnt]]
        i += cnt
        k += cnt
    tets, treg, tris, tbc = [], [], [], []
    E = sec["Elements"]
    nblocks = int(E[0].split()[0])
    i = 1
    for _ in range(nblocks):
        dim, ent, etype, cnt = (int(v) for v in E[i].split())
        i += 1
        ph = phys.get(dim, {}).get(ent, [])
        if etype == 4 and dim == 3 and ph:
            for ln in E[i:i + cnt]:
                tets.append([int(v) for v in ln.split()[1:5]]); treg.append(ph[0])
        elif etype == 2 and dim == 2 and ph:
            for ln in E[i:i + cnt]:
                tris.append([int(v) for v in ln.split()[1:4]]); tbc.append(ph[0])
        i += cnt
    return (tags, xyz, np.array(tets, dtype=np.int64).reshape(-1, 4), np.array(treg, dtype=np.int64),
            np.array(tris, dtype=np.int64).reshape(-1, 3), np.array(tbc, dtype=np.int64))


def write_msh(path, mesh: TetMesh, names=None):
    """MSH 2.2 ASCII (tets with physical volume = body id, triangles with physical surface = boundary id).
    ``names``: {(dim, id): "name"} -> ``$PhysicalNames`` (ids are then preserved by ``read_msh``)."""
    with open(path, "w") as f:
        f.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n")
        if names:
            f.write(f"$PhysicalNames\n{len(names)}\n")
            for (dim, pid), nm in sorted(names.items()):
                f.write(f'{dim} {pid} "{nm}"\n')
            f.write("$EndPhysicalNames\n")
        f.write(f"$Nodes\n{mesh.nn}\n")
        np.savetxt(f, np.column_stack([np.arange(1, mesh.nn + 1), mesh.nodes]), fmt=["%d", "%.17g", "%.17g", "%.17g"])
        f.write(f"$EndNodes\n$Elements\n{mesh.nb + mesh.nt}\n")
        nb, nt = mesh.nb, mesh.nt
        if nb:
            np.savetxt(f, np.column_stack([np.arange(1, nb + 1), np.full(nb, 2), np.full(nb, 2), mesh.bcid, mesh.bcid,
                                           mesh.tris + 1]), fmt="%d")
        np.savetxt(f, np.column_stack([np.arange(nb + 1, nb + nt + 1), np.full(nt, 4), np.full(nt, 2), mesh.region, mesh.region,
                                       mesh.tets + 1]), fmt="%d")
        f.write("$EndElements\n")


def elmergrid_14_2(msh_path, out_dir, renumber=None) -> TetMesh:
    """``ElmerGrid 14 2 <msh> -out <dir>``: Gmsh mesh -> Elmer mesh directory."""
    from . import elmer_io
    mesh = read_msh(msh_path, renumber=renumber)
    elmer_io.write_elmer_mesh(out_dir, mesh)
    return mesh
