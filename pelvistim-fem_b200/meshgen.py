"""Built-in tetrahedral meshers (host side, numpy).

The reference meshes with the ``gmsh`` Python API (OpenCASCADE + Delaunay):
``step01_box/box.geo:1-25``, ``step02_electrodes/run_sweep.py:55-130``,
``step03_ankle_layers/run_layered_sweep.py:122-362``.  gmsh is not available
where this engine runs, so the drivers fall back to these structured meshers,
which reproduce the same *geometry and physical tagging* (bodies 1..5,
boundaries 101/102/103) on a tensor-product grid whose hexahedra are split
into 6 Kuhn tetrahedra.  Meshes are P1 (4-node tets, 3-node boundary
triangles), node indices 0-based in memory (the Elmer mesh files are 1-based,
see ``elmer_io.py``).

Everything here is geometry only; no solver arithmetic.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# 6 Kuhn tets of the unit cube: paths 000 -> 111 along the 6 axis permutations.
# Corner code = bit0:x, bit1:y, bit2:z.
_KUHN = []
for _perm in ((0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)):
    _c = 0
    _path = [0]
    for _a in _perm:
        _c |= 1 << _a
        _path.append(_c)
    _KUHN.append(_path)
_KUHN = np.array(_KUHN, dtype=np.int64)            # [6, 4] corner codes


@dataclass
class TetMesh:
    """P1 tetrahedral mesh with tagged boundary triangles."""
    nodes: np.ndarray                # float64 [nn, 3]
    tets: np.ndarray                 # int32   [nt, 4]   0-based
    region: np.ndarray               # int32   [nt]      body id (1..)
    tris: np.ndarray                 # int32   [nb, 3]   0-based
    bcid: np.ndarray                 # int32   [nb]      boundary id
    tri_parent: np.ndarray | None = None   # int32 [nb] parent tet (0-based) or -1
    meta: dict = field(default_factory=dict)

    @property
    def nn(self):
        return int(self.nodes.shape[0])

    @property
    def nt(self):
        return int(self.tets.shape[0])

    @property
    def nb(self):
        return int(self.tris.shape[0])


def tet_volumes(nodes, tets):
    p = nodes[tets]
    a = p[:, 1] - p[:, 0]
    b = p[:, 2] - p[:, 0]
    c = p[:, 3] - p[:, 0]
    return np.einsum("ij,ij->i", a, np.cross(b, c)) / 6.0


def orient_positive(nodes, tets):
    """Swap two vertices of negatively oriented tets (in place) so that det > 0."""
    v = tet_volumes(nodes, tets)
    neg = v < 0
    if neg.any():
        t2 = tets[neg, 2].copy()
        tets[neg, 2] = tets[neg, 3]
        tets[neg, 3] = t2
    return tets


def _hex_to_tets(corner_ids):
    """corner_ids: int64 [nh, 8] (corner code order) -> tets [6*nh, 4]."""
    t = corner_ids[:, _KUHN]                       # [nh, 6, 4]
    return t.reshape(-1, 4)


def _grid_node_id(i, j, k, nxn, nyn):
    return (k * nyn + j) * nxn + i


def _face_tris(quad):
    """quad: [nq, 4] node ids ordered (00, 10, 11, 01) in face-local axes whose
    (1,1) corner is the Kuhn max corner -> the two triangles sharing the 00-11
    diagonal (the one the Kuhn split produces on every hex face)."""
    a = quad[:, [0, 1, 2]]
    b = quad[:, [0, 2, 3]]
    return np.concatenate([a, b], axis=0)


def find_tri_parents(tets, tris):
    """Parent tet of each boundary triangle (first tet containing the face) or -1.

    Vectorised face matching (sort node triples, searchsorted).  Intended for
    meshes up to a few million tets; the large synthetic meshes skip it.
    """
    nt = tets.shape[0]
    faces = np.concatenate([tets[:, [0, 1, 2]], tets[:, [0, 1, 3]],
                            tets[:, [0, 2, 3]], tets[:, [1, 2, 3]]], axis=0).astype(np.int64)
    faces.sort(axis=1)
    owner = np.tile(np.arange(nt, dtype=np.int64), 4)
    order = np.lexsort((owner, faces[:, 2], faces[:, 1], faces[:, 0]))
    fs = faces[order]
    ow = owner[order]
    q = np.sort(tris.astype(np.int64), axis=1)
    # pack sorted triples into a void view for searchsorted
    def pack(a):
        a = np.ascontiguousarray(a)
        return a.view([("a", np.int64), ("b", np.int64), ("c", np.int64)]).ravel()
    pos = np.searchsorted(pack(fs), pack(q), side="left")
    pos = np.minimum(pos, fs.shape[0] - 1)
    hit = (fs[pos] == q).all(axis=1)
    parent = np.where(hit, ow[pos], -1)
    return parent.astype(np.int32)


def external_faces(tets):
    """Faces that belong to exactly one tet -> (tris [nf,3] int32, parent [nf] int32).

    Same definition as ``step01_box/find_boundaries.py:48-59`` (a face seen once
    is a boundary face), vectorised.
    """
    nt = tets.shape[0]
    faces = np.concatenate([tets[:, [0, 1, 2]], tets[:, [0, 1, 3]],
                            tets[:, [0, 2, 3]], tets[:, [1, 2, 3]]], axis=0).astype(np.int64)
    owner = np.tile(np.arange(nt, dtype=np.int64), 4)
    key = np.sort(faces, axis=1)
    order = np.lexsort((key[:, 2], key[:, 1], key[:, 0]))
    ks = key[order]
    same_next = np.zeros(ks.shape[0], dtype=bool)
    same_next[:-1] = (ks[1:] == ks[:-1]).all(axis=1)
    same_prev = np.zeros(ks.shape[0], dtype=bool)
    same_prev[1:] = same_next[:-1]
    single = ~(same_next | same_prev)
    idx = np.sort(order[single])
    return faces[idx].astype(np.int32), owner[idx].astype(np.int32)


# ---------------------------------------------------------------------------
# step01: homogeneous box (box.geo)
# ---------------------------------------------------------------------------
def box_mesh(Lx=0.04, Ly=0.04, Lz=0.02, nx=10, ny=10, nz=5, jitter=0.0, seed=0,
             ids=(101, 102, 103), with_parents=True):
    """Box [0,Lx]x[0,Ly]x[0,Lz]; body 1; boundaries ids = (top, bottom, sides).

    Geometry and tags follow ``step01_box/box.geo:4-21`` (0.04x0.04x0.02 m,
    lc = 4 mm -> 10x10x5 cells; Physical Surface 101 top / 102 bottom / 103 sides).
    ``jitter`` moves interior nodes by U(-j*h, j*h) (seeded) to get a
    non-degenerate unstructured-like mesh for the analytic test.
    """
    xs = np.linspace(0.0, Lx, nx + 1)
    ys = np.linspace(0.0, Ly, ny + 1)
    zs = np.linspace(0.0, Lz, nz + 1)
    m = _tensor_mesh(xs, ys, zs, jitter=jitter, seed=seed)
    nxn, nyn, nzn = nx + 1, ny + 1, nz + 1
    top = _plane_quads(nxn, nyn, nzn, axis=2, index=nz)
    bot = _plane_quads(nxn, nyn, nzn, axis=2, index=0)
    sides = np.concatenate([_plane_quads(nxn, nyn, nzn, axis=0, index=0),
                            _plane_quads(nxn, nyn, nzn, axis=0, index=nx),
                            _plane_quads(nxn, nyn, nzn, axis=1, index=0),
                            _plane_quads(nxn, nyn, nzn, axis=1, index=ny)], axis=0)
    t_top, t_bot, t_side = _face_tris(top), _face_tris(bot), _face_tris(sides)
    tris = np.concatenate([t_top, t_bot, t_side], axis=0).astype(np.int32)
    bcid = np.concatenate([np.full(len(t_top), ids[0]), np.full(len(t_bot), ids[1]),
                           np.full(len(t_side), ids[2])]).astype(np.int32)
    region = np.ones(m["tets"].shape[0], dtype=np.int32)
    mesh = TetMesh(m["nodes"], m["tets"], region, tris, bcid,
                   meta=dict(kind="box", Lx=Lx, Ly=Ly, Lz=Lz, nx=nx, ny=ny, nz=nz))
    if with_parents:
        mesh.tri_parent = find_tri_parents(mesh.tets, mesh.tris)
    return mesh


def _tensor_mesh(xs, ys, zs, jitter=0.0, seed=0, frozen_z=None):
    """Kuhn-split tensor grid. Returns dict(nodes, tets, hex_ijk shape info)."""
    nxn, nyn, nzn = len(xs), len(ys), len(zs)
    X, Y, Z = np.meshgrid(xs, ys, zs, indexing="ij")          # [nxn, nyn, nzn]
    # node id = (k*nyn + j)*nxn + i  -> order axes (k, j, i)
    nodes = np.stack([X.transpose(2, 1, 0).ravel(), Y.transpose(2, 1, 0).ravel(),
                      Z.transpose(2, 1, 0).ravel()], axis=1)
    if jitter > 0.0:
        rng = np.random.default_rng(seed)
        hx = np.min(np.diff(xs)); hy = np.min(np.diff(ys)); hz = np.min(np.diff(zs))
        d = rng.uniform(-1.0, 1.0, size=nodes.shape) * jitter * np.array([hx, hy, hz])
        ii = np.tile(np.arange(nxn), nyn * nzn)
        jj = np.tile(np.repeat(np.arange(nyn), nxn), nzn)
        kk = np.repeat(np.arange(nzn), nxn * nyn)
        interior = (ii > 0) & (ii < nxn - 1) & (jj > 0) & (jj < nyn - 1) & (kk > 0) & (kk < nzn - 1)
        if frozen_z is not None:
            fz = np.zeros(nzn, dtype=bool)
            fz[list(frozen_z)] = True
            interior &= ~fz[kk]
        nodes[interior] += d[interior]
    i, j, k = np.meshgrid(np.arange(nxn - 1), np.arange(nyn - 1), np.arange(nzn - 1), indexing="ij")
    i = i.transpose(2, 1, 0).ravel(); j = j.transpose(2, 1, 0).ravel(); k = k.transpose(2, 1, 0).ravel()
    corners = np.empty((i.size, 8), dtype=np.int64)
    for c in range(8):
        corners[:, c] = _grid_node_id(i + (c & 1), j + ((c >> 1) & 1), k + ((c >> 2) & 1), nxn, nyn)
    tets = _hex_to_tets(corners).astype(np.int32)
    orient_positive(nodes, tets)
    return dict(nodes=nodes, tets=tets, hex_i=i, hex_j=j, hex_k=k, corners=corners)


def _plane_quads(nxn, nyn, nzn, axis, index, mask=None):
    """Quads (00,10,11,01 in the two in-plane axes, increasing) of the grid plane
    ``axis == index``.  ``mask`` optionally selects cells ([na-1, nb-1] bool, in
    (first in-plane axis, second in-plane axis) order)."""
    dims = [nxn, nyn, nzn]
    inplane = [a for a in range(3) if a != axis]
    na, nb = dims[inplane[0]], dims[inplane[1]]
    a, b = np.meshgrid(np.arange(na - 1), np.arange(nb - 1), indexing="ij")
    if mask is not None:
        a = a[mask]; b = b[mask]
    a = a.ravel(); b = b.ravel()

    def nid(da, db):
        idx = [None, None, None]
        idx[axis] = np.full(a.shape, index)
        idx[inplane[0]] = a + da
        idx[inplane[1]] = b + db
        return _grid_node_id(idx[0], idx[1], idx[2], nxn, nyn)
    return np.stack([nid(0, 0), nid(1, 0), nid(1, 1), nid(0, 1)], axis=1)


# ---------------------------------------------------------------------------
# step02: box with two electrode patches on the top face
# ---------------------------------------------------------------------------
def _footprint_mask(xc, yc, cx, cy, r, shape, match_area=True, cell_area=None):
    """Cells (centres xc,yc as 2-D arrays) inside a disk/square footprint.

    For disks the staircase set is chosen by centre distance; with
    ``match_area`` the number of cells is picked so that the staircase area is
    as close as possible to pi r^2 (area error below one cell)."""
    if shape == "square":
        return (np.abs(xc - cx) < r) & (np.abs(yc - cy) < r)
    d = np.hypot(xc - cx, yc - cy)
    inside = d < r
    if not match_area or cell_area is None:
        return inside
    target = np.pi * r * r
    order = np.argsort(d, axis=None, kind="stable")
    ca = cell_area.ravel()[order]
    cum = np.cumsum(ca)
    n = int(np.searchsorted(cum, target))
    if n < ca.size and n > 0 and abs(cum[n] - target) < abs(cum[n - 1] - target):
        n += 1
    n = max(n, 1)
    m = np.zeros(d.size, dtype=bool)
    m[order[:n]] = True
    return m.reshape(d.shape)


def _snap_rim_columns(nodes, nxn, nyn, nzn, xs, ys, masks, centers, r, max_shift=0.42):
    """Move the grid columns on the rim of each disk footprint radially onto the circle of radius ``r`` (the
    whole column, all z-levels), by at most ``max_shift`` of the local spacing, so the staircase outline of the
    structured grid follows the electrode edge.  Returns the displaced copy (topology unchanged)."""
    out = nodes.copy()
    hx = np.minimum(np.diff(xs, prepend=xs[0] - (xs[1] - xs[0])), np.diff(xs, append=xs[-1] + (xs[-1] - xs[-2])))
    hy = np.minimum(np.diff(ys, prepend=ys[0] - (ys[1] - ys[0])), np.diff(ys, append=ys[-1] + (ys[-1] - ys[-2])))
    for msk, (cx, cy) in zip(masks, centers):
        pad = np.zeros((nxn + 1, nyn + 1), dtype=bool)
        pad[1:nxn, 1:nyn] = msk                      # cell (i,j) occupies pad[i+1, j+1]
        # node (i,j) touches cells (i-1..i, j-1..j)
        touch_in = pad[0:nxn, 0:nyn] | pad[1:nxn + 1, 0:nyn] | pad[0:nxn, 1:nyn + 1] | pad[1:nxn + 1, 1:nyn + 1]
        touch_all = pad[0:nxn, 0:nyn] & pad[1:nxn + 1, 0:nyn] & pad[0:nxn, 1:nyn + 1] & pad[1:nxn + 1, 1:nyn + 1]
        rim = touch_in & ~touch_all
        ii, jj = np.nonzero(rim)
        interior = (ii > 0) & (ii < nxn - 1) & (jj > 0) & (jj < nyn - 1)
        ii, jj = ii[interior], jj[interior]
        dx, dy = xs[ii] - cx, ys[jj] - cy
        d = np.hypot(dx, dy)
        ok = d > 1e-12
        scale = np.where(ok, r / np.where(ok, d, 1.0), 1.0)
        sx, sy = dx * (scale - 1.0), dy * (scale - 1.0)
        lim = max_shift * np.minimum(hx[ii], hy[jj])
        mag = np.hypot(sx, sy)
        f = np.where(mag > lim, lim / np.where(mag > 0, mag, 1.0), 1.0)
        sx, sy = sx * f, sy * f
        for k in range(nzn):
            nid = _grid_node_id(ii, jj, k, nxn, nyn)
            out[nid, 0] += sx
            out[nid, 1] += sy
    return out


def graded_lines(L, h_far, refine=None):
    """1-D grid lines on [0, L]: spacing ~h_far, refined to ~h_near inside the
    intervals listed in ``refine`` = [(lo, hi, h_near), ...] (merged, clipped)."""
    if not refine:
        n = max(1, int(round(L / h_far)))
        return np.linspace(0.0, L, n + 1)
    # piecewise-constant target spacing, integrate to get monotone map
    pts = {0.0, float(L)}
    for lo, hi, _ in refine:
        pts.add(min(max(lo, 0.0), L)); pts.add(min(max(hi, 0.0), L))
    pts = sorted(pts)
    lines = [0.0]
    for a, b in zip(pts[:-1], pts[1:]):
        if b - a < 1e-12:
            continue
        mid = 0.5 * (a + b)
        h = h_far
        for lo, hi, hn in refine:
            if lo - 1e-12 <= mid <= hi + 1e-12:
                h = min(h, hn)
        n = max(1, int(np.ceil((b - a) / h - 1e-9)))
        seg = np.linspace(a, b, n + 1)[1:]
        lines.extend(seg.tolist())
    return np.array(lines)


def electrode_box_mesh(Lx, Ly, Lz, e1_xy, e2_xy, r, shape="circle", h_elec=None, h_bulk=None,
                       nz=None, with_parents=True, snap_rim=False):
    """Homogeneous box with two electrode patches on the top face.

    Geometry/tags follow ``step02_electrodes/run_sweep.py:39-52,63-103``: box
    Lx x Ly x Lz, patches (disk radius r / square half-side r) on z = Lz centred
    at e1_xy (active, 101) and e2_xy (return, 102); everything else 103;
    one body (1).  Element size ~ r/3.5 near the patches, min(4r, 12 mm) away
    (``run_sweep.py:109-110``).
    """
    h_elec = h_elec if h_elec is not None else r / 3.5
    h_bulk = h_bulk if h_bulk is not None else min(4 * r, 0.012)
    pad = 1.5 * r
    xs = graded_lines(Lx, h_bulk, [(e1_xy[0] - pad, e1_xy[0] + pad, h_elec),
                                   (e2_xy[0] - pad, e2_xy[0] + pad, h_elec)])
    ys = graded_lines(Ly, h_bulk, [(min(e1_xy[1], e2_xy[1]) - pad, max(e1_xy[1], e2_xy[1]) + pad, h_elec)])
    if nz is None:
        # geometric grading in z: fine at the top (electrodes), coarse at the bottom
        zs_rev = [0.0]
        h = h_elec
        while zs_rev[-1] < Lz - 1e-12:
            zs_rev.append(min(Lz, zs_rev[-1] + h))
            h = min(h * 1.3, h_bulk)
        zs = Lz - np.array(zs_rev[::-1])
        zs[0] = 0.0; zs[-1] = Lz
    else:
        zs = np.linspace(0.0, Lz, nz + 1)
    m = _tensor_mesh(xs, ys, zs)
    nxn, nyn, nzn = len(xs), len(ys), len(zs)
    xc = 0.5 * (xs[:-1] + xs[1:])[:, None] * np.ones((1, nyn - 1))
    yc = np.ones((nxn - 1, 1)) * 0.5 * (ys[:-1] + ys[1:])[None, :]
    ca = np.diff(xs)[:, None] * np.diff(ys)[None, :]
    m1 = _footprint_mask(xc, yc, e1_xy[0], e1_xy[1], r, shape, cell_area=ca)
    m2 = _footprint_mask(xc, yc, e2_xy[0], e2_xy[1], r, shape, cell_area=ca)
    if snap_rim and shape == "circle":
        shift = 0.42
        while shift > 0.05:
            cand = _snap_rim_columns(m["nodes"], nxn, nyn, nzn, xs, ys, (m1, m2), (e1_xy, e2_xy), r, shift)
            if (tet_volumes(cand, m["tets"]) > 0).all():
                m["nodes"] = cand
                break
            shift *= 0.5
    rest = ~(m1 | m2)
    k_top = nzn - 1
    q1 = _plane_quads(nxn, nyn, nzn, 2, k_top, m1)
    q2 = _plane_quads(nxn, nyn, nzn, 2, k_top, m2)
    q3 = np.concatenate([_plane_quads(nxn, nyn, nzn, 2, k_top, rest),
                         _plane_quads(nxn, nyn, nzn, 2, 0),
                         _plane_quads(nxn, nyn, nzn, 0, 0), _plane_quads(nxn, nyn, nzn, 0, nxn - 1),
                         _plane_quads(nxn, nyn, nzn, 1, 0), _plane_quads(nxn, nyn, nzn, 1, nyn - 1)], axis=0)
    t1, t2, t3 = _face_tris(q1), _face_tris(q2), _face_tris(q3)
    tris = np.concatenate([t1, t2, t3], axis=0).astype(np.int32)
    bcid = np.concatenate([np.full(len(t1), 101), np.full(len(t2), 102), np.full(len(t3), 103)]).astype(np.int32)
    mesh = TetMesh(m["nodes"], m["tets"], np.ones(m["tets"].shape[0], dtype=np.int32), tris, bcid,
                   meta=dict(kind="electrode_box", Lx=Lx, Ly=Ly, Lz=Lz, r=r, shape=shape))
    if with_parents:
        mesh.tri_parent = find_tri_parents(mesh.tets, mesh.tris)
    return mesh


# ---------------------------------------------------------------------------
# step03 / step04 / synthetic: layered slab with contact pads
# ---------------------------------------------------------------------------
def layered_slab_mesh(Lx=0.080, Ly=0.060, Lz=0.040, t_skin=0.0015, t_fat=0.005, t_contact=0.0005,
                      active_xy=(0.015, 0.045), return_xy=(0.065, 0.045), elec_r=0.010, shape="circle",
                      xs=None, ys=None, n_muscle=12, n_fat=3, n_skin=2, n_contact=1,
                      h_bulk=0.003, h_elec=0.0015, jitter=0.0, seed=0,
                      interfaces_as_103=True, with_parents=True, contact_enabled=True, snap_rim=False, bone=None):
    """Layered slab: muscle (body 1) / fat (2) / skin (3) + two contact pads
    (4 active, 5 return) sitting ON TOP of the skin.

    Geometry/tags follow ``step03_ankle_layers/run_layered_sweep.py:142-181,206-227,
    300-308`` (rect cross-section): slab [0,Lx]x[0,Ly]x[0,Lz]; muscle
    z in [0, Lz-t_skin-t_fat], fat above, skin on top; pads z in [Lz, Lz+t_contact]
    with disk (radius r) or square (half-side r) footprint; boundary 101 = active
    pad top face, 102 = return pad top face, 103 = every other 2-D entity,
    including the internal layer interfaces (``other_s``, ``:297,308``) when
    ``interfaces_as_103``.  Without contact (``contact_enabled=False``) the
    electrode patches lie on the skin top face.

    ``bone``: optional ``dict(x=(x0, x1), y=(y0, y1), z=(z0, z1))`` - a block of body 6 (bone; BASELINE.json north_star
    "skin/fat/muscle/bone layers"; the reference's slab only names the bone faces, ``params.yaml:9-10,18``) cut out of the
    muscle.  Its faces are snapped onto the nearest grid planes; the extents actually meshed are in ``meta["bone"]``.
    """
    t_muscle = Lz - t_skin - t_fat
    if t_muscle <= 1e-4:
        raise ValueError(f"t_muscle = {t_muscle*1000:.2f} mm <= 0.1 mm - reduce t_fat + t_skin or increase Lz")
    pad = 1.5 * elec_r
    if xs is None:
        xs = graded_lines(Lx, h_bulk, [(active_xy[0] - pad, active_xy[0] + pad, h_elec),
                                       (return_xy[0] - pad, return_xy[0] + pad, h_elec)])
    if ys is None:
        ys = graded_lines(Ly, h_bulk, [(min(active_xy[1], return_xy[1]) - pad,
                                        max(active_xy[1], return_xy[1]) + pad, h_elec)])
    xs = np.asarray(xs, dtype=np.float64); ys = np.asarray(ys, dtype=np.float64)
    z0_fat = t_muscle
    z0_skin = t_muscle + t_fat
    zl = [np.linspace(0.0, z0_fat, n_muscle + 1),
          np.linspace(z0_fat, z0_skin, n_fat + 1)[1:],
          np.linspace(z0_skin, Lz, n_skin + 1)[1:]]
    nc = n_contact if contact_enabled else 0
    if nc:
        zl.append(np.linspace(Lz, Lz + t_contact, nc + 1)[1:])
    zs = np.concatenate(zl)
    zs[n_muscle] = z0_fat; zs[n_muscle + n_fat] = z0_skin; zs[n_muscle + n_fat + n_skin] = Lz
    k_fat, k_skin, k_top = n_muscle, n_muscle + n_fat, n_muscle + n_fat + n_skin
    frozen = {k_fat, k_skin, k_top} | set(range(k_top, len(zs)))
    m = _tensor_mesh(xs, ys, zs, jitter=jitter, seed=seed, frozen_z=frozen)
    nxn, nyn, nzn = len(xs), len(ys), len(zs)
    hi, hj, hk = m["hex_i"], m["hex_j"], m["hex_k"]
    xc2 = 0.5 * (xs[:-1] + xs[1:])[:, None] * np.ones((1, nyn - 1))
    yc2 = np.ones((nxn - 1, 1)) * 0.5 * (ys[:-1] + ys[1:])[None, :]
    ca = np.diff(xs)[:, None] * np.diff(ys)[None, :]
    m1 = _footprint_mask(xc2, yc2, active_xy[0], active_xy[1], elec_r, shape, cell_area=ca)
    m2 = _footprint_mask(xc2, yc2, return_xy[0], return_xy[1], elec_r, shape, cell_area=ca)
    if (m1 & m2).any():
        raise ValueError("electrode footprints overlap")
    # hex body ids
    body = np.where(hk < k_fat, 1, np.where(hk < k_skin, 2, 3)).astype(np.int32)
    bone_meta = None
    if bone is not None:
        def snap(lines, lo, hi, kmax):
            a = int(np.argmin(np.abs(lines[:kmax + 1] - lo)))
            b = int(np.argmin(np.abs(lines[:kmax + 1] - hi)))
            if b <= a:
                raise ValueError("bone block is thinner than one cell of this mesh")
            return a, b
        ia, ib = snap(xs, bone["x"][0], bone["x"][1], nxn - 1)
        ja, jb = snap(ys, bone["y"][0], bone["y"][1], nyn - 1)
        ka, kb = snap(zs, bone["z"][0], bone["z"][1], k_fat)          # bone lies inside the muscle block
        inb = (hi >= ia) & (hi < ib) & (hj >= ja) & (hj < jb) & (hk >= ka) & (hk < kb)
        body = np.where(inb, 6, body).astype(np.int32)
        bone_meta = dict(x=(float(xs[ia]), float(xs[ib])), y=(float(ys[ja]), float(ys[jb])), z=(float(zs[ka]), float(zs[kb])))
    keep = np.ones(hk.shape, dtype=bool)
    if nc:
        in_pad = hk >= k_top
        a = m1[hi, hj]; b = m2[hi, hj]
        body = np.where(in_pad & a, 4, body)
        body = np.where(in_pad & b, 5, body)
        keep = ~in_pad | a | b
    corners = m["corners"][keep]
    body = body[keep]
    tets = _hex_to_tets(corners).astype(np.int32)
    region = np.repeat(body, 6).astype(np.int32)
    nodes = m["nodes"]
    orient_positive(nodes, tets)
    if snap_rim and shape == "circle":
        # follow the circular electrode edge instead of the grid's staircase (halved until no element inverts)
        shift = 0.42
        while shift > 0.05:
            cand = _snap_rim_columns(nodes, nxn, nyn, nzn, xs, ys, (m1, m2), (active_xy, return_xy), elec_r, shift)
            if (tet_volumes(cand, tets) > 0).all():
                nodes = cand
                break
            shift *= 0.5
    # boundary triangles
    k_elec = nzn - 1 if nc else k_top
    q101 = _plane_quads(nxn, nyn, nzn, 2, k_elec, m1)
    q102 = _plane_quads(nxn, nyn, nzn, 2, k_elec, m2)
    others = []
    rest = ~(m1 | m2)
    if nc:
        others.append(_plane_quads(nxn, nyn, nzn, 2, k_top, rest))          # exposed skin top
        # pad side walls: footprint cell edges whose neighbour is outside the footprint
        for msk in (m1, m2):
            for kk in range(k_top, nzn - 1):
                others.append(_pad_wall_quads(msk, nxn, nyn, kk))
        if interfaces_as_103:
            others.append(_plane_quads(nxn, nyn, nzn, 2, k_top, m1 | m2))    # skin/pad interface
    else:
        others.append(_plane_quads(nxn, nyn, nzn, 2, k_top, rest))
    others.append(_plane_quads(nxn, nyn, nzn, 2, 0))
    if interfaces_as_103:
        others.append(_plane_quads(nxn, nyn, nzn, 2, k_fat))
        others.append(_plane_quads(nxn, nyn, nzn, 2, k_skin))
    # outer side walls of the slab (k < k_top)
    for axis, index in ((0, 0), (0, nxn - 1), (1, 0), (1, nyn - 1)):
        dims = [nxn, nyn, nzn]
        inpl = [a_ for a_ in range(3) if a_ != axis]
        msk = np.zeros((dims[inpl[0]] - 1, dims[inpl[1]] - 1), dtype=bool)
        msk[:, :k_top] = True
        others.append(_plane_quads(nxn, nyn, nzn, axis, index, msk))
    q103 = np.concatenate(others, axis=0)
    t1, t2, t3 = _face_tris(q101), _face_tris(q102), _face_tris(q103)
    tris = np.concatenate([t1, t2, t3], axis=0)
    bcid = np.concatenate([np.full(len(t1), 101), np.full(len(t2), 102), np.full(len(t3), 103)]).astype(np.int32)
    # compact node numbering (unused pad-layer nodes are dropped)
    used = np.zeros(nodes.shape[0], dtype=bool)
    used[tets.ravel()] = True
    if not used.all():
        new_id = np.cumsum(used) - 1
        nodes = nodes[used]
        tets = new_id[tets].astype(np.int32)
        tris = new_id[tris]
    z_elec_top = Lz + (t_contact if nc else 0.0)
    mesh = TetMesh(np.ascontiguousarray(nodes), np.ascontiguousarray(tets), region,
                   np.ascontiguousarray(tris.astype(np.int32)), bcid,
                   meta=dict(kind="layered_slab", Lx=Lx, Ly=Ly, Lz=Lz, t_skin=t_skin, t_fat=t_fat,
                             t_contact=t_contact if nc else 0.0, elec_r=elec_r, shape=shape,
                             active_xy=tuple(active_xy), return_xy=tuple(return_xy),
                             z_elec_top=z_elec_top, contact_enabled=bool(nc),
                             bone=bone_meta,
                             area_active=float(ca[m1].sum()), area_return=float(ca[m2].sum()),
                             grid=(nxn, nyn, nzn)))
    if with_parents:
        mesh.tri_parent = find_tri_parents(mesh.tets, mesh.tris)
    return mesh


def _pad_wall_quads(msk, nxn, nyn, k):
    """Vertical quads on the rim of footprint ``msk`` between z-levels k and k+1."""
    quads = []
    nxc, nyc = msk.shape
    padm = np.zeros((nxc + 2, nyc + 2), dtype=bool)
    padm[1:-1, 1:-1] = msk
    # x-normal walls: cell (i,j) in, neighbour (i-1,j) out  -> wall at x-index i ; (i+1,j) out -> wall at i+1
    for di, off in ((-1, 0), (1, 1)):
        wall = msk & ~padm[1 + di:nxc + 1 + di, 1:-1]
        i, j = np.nonzero(wall)
        xi = i + off
        n00 = _grid_node_id(xi, j, k, nxn, nyn); n10 = _grid_node_id(xi, j + 1, k, nxn, nyn)
        n11 = _grid_node_id(xi, j + 1, k + 1, nxn, nyn); n01 = _grid_node_id(xi, j, k + 1, nxn, nyn)
        quads.append(np.stack([n00, n10, n11, n01], axis=1))
    for dj, off in ((-1, 0), (1, 1)):
        wall = msk & ~padm[1:-1, 1 + dj:nyc + 1 + dj]
        i, j = np.nonzero(wall)
        yj = j + off
        n00 = _grid_node_id(i, yj, k, nxn, nyn); n10 = _grid_node_id(i + 1, yj, k, nxn, nyn)
        n11 = _grid_node_id(i + 1, yj, k + 1, nxn, nyn); n01 = _grid_node_id(i, yj, k + 1, nxn, nyn)
        quads.append(np.stack([n00, n10, n11, n01], axis=1))
    return np.concatenate(quads, axis=0) if quads else np.zeros((0, 4), dtype=np.int64)


def delaunay_box_mesh(npts=2000, Lx=0.04, Ly=0.03, Lz=0.02, seed=0, ids=(101, 102, 103), min_quality=1e-9):
    """Unstructured tetrahedral mesh of a box: Delaunay triangulation (scipy / Qhull) of random interior points
    plus a jittered lattice on the faces, the kind of irregular connectivity (3..40 neighbours per node, no
    row-to-row coherence) a Gmsh mesh has.  Boundary ids = (top z=Lz, bottom z=0, sides).  Test meshes only."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    nface = max(4, int(round((npts / 6) ** (1 / 3) * 2)))
    g = np.linspace(0.0, 1.0, nface)
    faces = []
    for ax in range(3):
        for v in (0.0, 1.0):
            a, b = np.meshgrid(g, g, indexing="ij")
            a = a + rng.uniform(-0.3, 0.3, a.shape) / (nface - 1) * ((a > 0) & (a < 1))
            b = b + rng.uniform(-0.3, 0.3, b.shape) / (nface - 1) * ((b > 0) & (b < 1))
            p = np.empty((a.size, 3))
            p[:, ax] = v
            p[:, (ax + 1) % 3] = a.ravel()
            p[:, (ax + 2) % 3] = b.ravel()
            faces.append(p)
    surf = np.unique(np.round(np.concatenate(faces), 12), axis=0)
    inner = rng.uniform(0.03, 0.97, size=(max(npts - surf.shape[0], 8), 3))
    pts = np.concatenate([surf, inner]) * np.array([Lx, Ly, Lz])
    tets = Delaunay(pts).simplices.astype(np.int32)
    vol = tet_volumes(pts, tets)
    h3 = (Lx * Ly * Lz) / max(tets.shape[0], 1)
    tets = tets[np.abs(vol) > min_quality * h3]                 # drop the exactly flat tets Qhull leaves on the faces
    orient_positive(pts, tets)
    used = np.zeros(pts.shape[0], dtype=bool)
    used[tets.ravel()] = True
    new_id = np.cumsum(used) - 1
    pts, tets = pts[used], new_id[tets].astype(np.int32)
    tris, parent = external_faces(tets)
    zc = pts[tris][:, :, 2]
    bcid = np.where(np.all(zc > Lz * (1 - 1e-9), axis=1), ids[0], np.where(np.all(zc < Lz * 1e-9, axis=1), ids[1], ids[2])).astype(np.int32)
    return TetMesh(np.ascontiguousarray(pts), np.ascontiguousarray(tets), np.ones(tets.shape[0], dtype=np.int32),
                   np.ascontiguousarray(tris), bcid, tri_parent=parent, meta=dict(kind="delaunay_box", Lx=Lx, Ly=Ly, Lz=Lz))


def compress_under_pads(nodes, pads_xy, pad_r, depth, Lz, width_factor=1.5):
    """Node displacement of tissue compressed by the electrodes: a Gaussian indentation of ``depth`` (m) centred
    on each pad, growing linearly from 0 at z = 0 to its full value at the skin surface (pads ride down
    rigidly).  Same idea as the reference's Gaussian skin-height field applied to fixed topology
    (``step03_ankle_layers/run_layered_sweep.py:93-118,329-340``); used with ``ptfem_mesh_set_coords``.
    The indentation is clipped to 60 % of the local column height so no element inverts."""
    out = np.array(nodes, dtype=np.float64, copy=True)
    sig2 = (width_factor * pad_r) ** 2
    dz = np.zeros(out.shape[0])
    for (px, py) in pads_xy:
        dz += depth * np.exp(-((out[:, 0] - px) ** 2 + (out[:, 1] - py) ** 2) / (2.0 * sig2))
    dz = np.minimum(dz, 0.6 * Lz)
    out[:, 2] -= dz * np.clip(out[:, 2] / Lz, 0.0, 1.0)
    return out


# Synthetic benchmark meshes: uniform tensor grid on the
# step03 slab, z-levels snapped to the layer interfaces.
SYNTH_SIZES = {
    "XS": (16, 12, 10),
    "S": (32, 24, 16),
    "M": (96, 72, 60),
    "L": (192, 144, 120),
    "XL": (288, 216, 180),     # 67 M tets, 11.3 M nodes: the row-partitioned solve with a distributed set-up
}


def synth_slab(size="S", seed=0, jitter=0.2, elec_r=0.010, with_parents=False, interfaces_as_103=False,
               contact_enabled=True):
    """Synthetic refined layered slab of SURVEY.md section 8(d): 80x60x40.5 mm, nx*ny*nz
    hexes -> 6 Kuhn tets each, regions 1-5, pads r = 10 mm at (15,45)/(65,45) mm."""
    nx, ny, nz = SYNTH_SIZES[size] if isinstance(size, str) else size
    Lx, Ly, Lz, t_skin, t_fat, t_c = 0.080, 0.060, 0.040, 0.0015, 0.005, 0.0005
    t_m = Lz - t_skin - t_fat
    # distribute nz cells over (muscle, fat, skin, pad) ~ proportional to thickness
    n_c = 1
    n_s = max(2, int(round(nz * t_skin / (Lz + t_c))))
    n_f = max(2, int(round(nz * t_fat / (Lz + t_c))))
    n_m = max(2, nz - n_c - n_s - n_f)
    xs = np.linspace(0.0, Lx, nx + 1)
    ys = np.linspace(0.0, Ly, ny + 1)
    return layered_slab_mesh(Lx, Ly, Lz, t_skin, t_fat, t_c, (0.015, 0.045), (0.065, 0.045), elec_r, "circle",
                             xs=xs, ys=ys, n_muscle=n_m, n_fat=n_f, n_skin=n_s, n_contact=n_c,
                             jitter=jitter, seed=seed, interfaces_as_103=interfaces_as_103,
                             with_parents=with_parents, contact_enabled=contact_enabled)


def tri_areas(nodes, tris):
    p = nodes[tris]
    return 0.5 * np.linalg.norm(np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]), axis=1)
