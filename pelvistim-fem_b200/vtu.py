"""VTK unstructured-grid (``.vtu``) writer / reader for the solve step's output.

The reference's solver writes ``results/case_t0001.vtu`` (``Output File Name =
"case"``, ``Output Format = VTU``, ``Save Geometry IDs = True``:
``step01_box/case.sif:47-54``) and every consumer reads it with pyvista:
points in mesh node order, cells = tetrahedra (VTK type 10) followed by the
boundary triangles (type 5, relied on by
``step03_ankle_layers/run_layered_sweep.py:716-725``), point data ``potential``
and ``volume current`` (3-vector), cell data ``GeometryIds``.

Written as XML with raw appended binary blocks (UInt64 headers, little endian),
which VTK/pyvista read natively; ``read_vtu`` reads files written here.
"""
from __future__ import annotations

import re
import struct
from pathlib import Path

import numpy as np

_VTK_TYPE = {np.dtype("float64"): "Float64", np.dtype("float32"): "Float32", np.dtype("int32"): "Int32",
             np.dtype("int64"): "Int64", np.dtype("uint8"): "UInt8"}
_NP_TYPE = {v: k for k, v in _VTK_TYPE.items()}


def write_vtu(path, nodes, tets, tris, point_data=None, cell_data=None):
    """Write points + (tets, tris) cells with named point / cell arrays."""
    nodes = np.ascontiguousarray(nodes, dtype=np.float64)
    tets = np.ascontiguousarray(tets, dtype=np.int32)
    tris = np.ascontiguousarray(tris, dtype=np.int32)
    nt, nb = tets.shape[0], tris.shape[0]
    conn = np.concatenate([tets.ravel(), tris.ravel()]).astype(np.int32)
    offsets = np.concatenate([4 * np.arange(1, nt + 1), 4 * nt + 3 * np.arange(1, nb + 1)]).astype(np.int32)
    types = np.concatenate([np.full(nt, 10), np.full(nb, 5)]).astype(np.uint8)
    blocks = []
    off = 0

    def darray(name, arr, ncomp=None):
        nonlocal off
        arr = np.ascontiguousarray(arr)
        tname = _VTK_TYPE[arr.dtype]
        nc = f' NumberOfComponents="{ncomp}"' if ncomp else ""
        nm = f' Name="{name}"' if name else ""
        s = f'<DataArray type="{tname}"{nm}{nc} format="appended" offset="{off}"/>'
        blocks.append(arr)
        off += 8 + arr.nbytes
        return s

    lines = ['<?xml version="1.0"?>',
             '<VTKFile type="UnstructuredGrid" version="1.0" byte_order="LittleEndian" header_type="UInt64">',
             '<UnstructuredGrid>',
             f'<Piece NumberOfPoints="{nodes.shape[0]}" NumberOfCells="{nt + nb}">']
    lines.append("<PointData>")
    for name, arr in (point_data or {}).items():
        arr = np.asarray(arr)
        lines.append(darray(name, arr, arr.shape[1] if arr.ndim == 2 else None))
    lines.append("</PointData>")
    lines.append("<CellData>")
    for name, arr in (cell_data or {}).items():
        arr = np.asarray(arr)
        lines.append(darray(name, arr, arr.shape[1] if arr.ndim == 2 else None))
    lines.append("</CellData>")
    lines.append("<Points>")
    lines.append(darray(None, nodes, 3))
    lines.append("</Points>")
    lines.append("<Cells>")
    lines.append(darray("connectivity", conn))
    lines.append(darray("offsets", offsets))
    lines.append(darray("types", types))
    lines.append("</Cells>")
    lines += ["</Piece>", "</UnstructuredGrid>", '<AppendedData encoding="raw">']
    with open(path, "wb") as f:
        f.write(("\n".join(lines) + "\n_").encode())
        for b in blocks:
            f.write(struct.pack("<Q", b.nbytes))
            f.write(b.tobytes())
        f.write(b"\n</AppendedData>\n</VTKFile>\n")


def read_vtu(path):
    """Read a VTU written by :func:`write_vtu`.

    Returns dict(points, cells_conn, cells_off, cell_types, point_data{}, cell_data{})."""
    raw = Path(path).read_bytes()
    m = re.search(rb'<AppendedData encoding="raw">\s*_', raw)
    if not m:
        raise ValueError("only raw appended VTU files are supported by this reader")
    head = raw[:m.start()].decode()
    base = m.end()

    def load(tag_text):
        t = re.search(r'type="(\w+)"', tag_text).group(1)
        o = int(re.search(r'offset="(\d+)"', tag_text).group(1))
        nc = re.search(r'NumberOfComponents="(\d+)"', tag_text)
        nbytes = struct.unpack_from("<Q", raw, base + o)[0]
        a = np.frombuffer(raw, dtype=_NP_TYPE[t], count=nbytes // _NP_TYPE[t].itemsize, offset=base + o + 8).copy()
        if nc and int(nc.group(1)) > 1:
            a = a.reshape(-1, int(nc.group(1)))
        return a

    def section(name):
        s = re.search(rf"<{name}>(.*?)</{name}>", head, re.S)
        out = {}
        if s:
            for tag in re.findall(r"<DataArray[^>]*/>", s.group(1)):
                nm = re.search(r'Name="([^"]*)"', tag)
                out[nm.group(1) if nm else ""] = load(tag)
        return out

    pts = section("Points")[""]
    cells = section("Cells")
    return dict(points=pts, cells_conn=cells["connectivity"], cells_off=cells["offsets"],
                cell_types=cells["types"], point_data=section("PointData"), cell_data=section("CellData"))


def split_cells(v):
    """(tets [nt,4], tris [nb,3]) from a :func:`read_vtu` result."""
    types = v["cell_types"]
    off = np.concatenate([[0], v["cells_off"]]).astype(np.int64)
    conn = v["cells_conn"]
    t_idx = np.nonzero(types == 10)[0]
    b_idx = np.nonzero(types == 5)[0]
    tets = conn[off[t_idx][:, None] + np.arange(4)] if t_idx.size else np.zeros((0, 4), np.int32)
    tris = conn[off[b_idx][:, None] + np.arange(3)] if b_idx.size else np.zeros((0, 3), np.int32)
    return tets, tris
