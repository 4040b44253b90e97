"""Graded unstructured mesher driven by a size field (host side, numpy + scipy.spatial).

The reference meshes the layered slab with Gmsh: OpenCASCADE volumes, a ``Distance`` field to the two electrode
surfaces fed into a ``Threshold`` field (``SizeMin = lc_electrode`` up to ``DistMin = elec_r``, growing linearly to
``SizeMax = lc_global`` at ``DistMax = 6 elec_r``; ``step03_ankle_layers/run_layered_sweep.py:311-323``, sizes in
``params.yaml:67-70``), so the pad rim is a polygon inscribed in the circle with ``round(2 pi r / lc_electrode)`` vertices
and the elements grow away from the pads.  Gmsh is not available where this engine runs; this module builds a mesh of the
same kind without it:

1. a 2-D triangulation of the slab's footprint whose edge length follows the same Distance/Threshold law - points from a
   thinned hexagonal lattice relaxed by the force-equilibrium iteration of Persson & Strang ("A simple mesh generator in
   MATLAB", SIAM Review 46, 2004) with Delaunay re-triangulation; the rim vertices of both pads are fixed points, and their
   polygon edges are guaranteed to be edges of the triangulation (no free point is left inside an edge's diametral
   circle), so the footprint is conforming;
2. extrusion of the triangulation through the layer stack (muscle / fat / skin levels, pad levels over the footprints
   only), level spacing in the muscle growing with the distance from the pads by the same law;
3. every prism split into three tetrahedra, the diagonal of each quadrilateral face running from the bottom of the
   lower-numbered column to the top of the higher-numbered one, which makes neighbouring prisms (and the boundary
   triangles on walls) agree without any search.

Tags are the reference's (bodies 1..5, bone 6; boundaries 101/102/103 with interfaces counted as 103,
``run_layered_sweep.py:296-308``).  Geometry only; no solver arithmetic.
"""
from __future__ import annotations

import numpy as np

from .meshgen import TetMesh, find_tri_parents, orient_positive, tet_volumes

_SQ3 = np.sqrt(3.0)


# ---------------------------------------------------------------------------
# size field (Distance -> Threshold, run_layered_sweep.py:311-320)
# ---------------------------------------------------------------------------
def pad_distance_2d(p, centers, r, shape="circle"):
    """Distance in the plane from points ``p`` [n,2] to the nearest pad footprint (0 inside)."""
    d = np.full(p.shape[0], np.inf)
    for cx, cy in centers:
        if shape == "square":
            qx = np.maximum(np.abs(p[:, 0] - cx) - r, 0.0)
            qy = np.maximum(np.abs(p[:, 1] - cy) - r, 0.0)
            d = np.minimum(d, np.hypot(qx, qy))
        else:
            d = np.minimum(d, np.maximum(np.hypot(p[:, 0] - cx, p[:, 1] - cy) - r, 0.0))
    return d


def threshold(dist, lc_min, lc_max, dist_min, dist_max):
    """Gmsh's Threshold law: lc_min below dist_min, lc_max above dist_max, linear in between."""
    t = np.clip((np.asarray(dist, dtype=np.float64) - dist_min) / max(dist_max - dist_min, 1e-300), 0.0, 1.0)
    return lc_min + t * (lc_max - lc_min)


# ---------------------------------------------------------------------------
# 2-D graded triangulation of a rectangle with fixed polygon loops
# ---------------------------------------------------------------------------
def _edges_of(tri):
    e = np.concatenate([tri[:, [0, 1]], tri[:, [1, 2]], tri[:, [2, 0]]], axis=0)
    e.sort(axis=1)
    return np.unique(e, axis=0)


def _tri_area(p, tri):
    a, b, c = p[tri[:, 0]], p[tri[:, 1]], p[tri[:, 2]]
    return 0.5 * ((b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0]))


def _delaunay(p, h0):
    from scipy.spatial import Delaunay
    tri = Delaunay(p).simplices.astype(np.int64)
    ar = _tri_area(p, tri)
    flip = ar < 0
    tri[flip] = tri[flip][:, [0, 2, 1]]
    return tri[np.abs(ar) > 1e-9 * h0 * h0]           # collinear boundary triples carry no area


def pad_loop(center, r, lc, shape="circle", Lx=None, Ly=None):
    """Rim vertices of one pad, counter-clockwise: circle -> inscribed polygon of about 2 pi r / lc vertices (what a 1-D
    mesh of the circle at size lc gives); square -> its edges cut into round(2 r / lc) pieces.  Where the circle touches a
    slab edge (the largest pads of the sweep are tangent to two of them, ``params.yaml:56-60``) the tangent point is a
    vertex - a CAD kernel splits the circle there - and each arc between tangent points is divided on its own."""
    cx, cy = center
    if shape == "square":
        n = max(1, int(round(2 * r / lc)))
        s = np.linspace(-r, r, n + 1)[:-1]
        pts = ([(cx + v, cy - r) for v in s] + [(cx + r, cy + v) for v in s] +
               [(cx - v, cy + r) for v in s] + [(cx - r, cy - v) for v in s])
        return np.array(pts)
    tol = 1e-6 * r
    cuts = []
    if Lx is not None and Ly is not None:
        if min(cx - r, cy - r, Lx - cx - r, Ly - cy - r) < -tol:
            raise ValueError("electrode footprint leaves the slab")
        for hit, ang in ((abs(Lx - cx - r) < tol, 0.0), (abs(Ly - cy - r) < tol, 0.5 * np.pi),
                         (abs(cx - r) < tol, np.pi), (abs(cy - r) < tol, 1.5 * np.pi)):
            if hit:
                cuts.append(ang)
    if not cuts:
        n = max(6, int(round(2 * np.pi * r / lc)))
        a = 2 * np.pi * np.arange(n) / n
    else:
        a = []
        for a0, a1 in zip(cuts, cuts[1:] + [cuts[0] + 2 * np.pi]):
            n = max(2, int(round((a1 - a0) * r / lc)))
            a.append(a0 + (a1 - a0) * np.arange(n) / n)
        a = np.concatenate(a)
    pts = np.stack([cx + r * np.cos(a), cy + r * np.sin(a)], axis=1)
    if cuts:                                             # tangent vertices lie exactly on the slab edge
        pts[:, 0] = np.where(np.abs(pts[:, 0]) < tol, 0.0, np.where(np.abs(pts[:, 0] - Lx) < tol, Lx, pts[:, 0]))
        pts[:, 1] = np.where(np.abs(pts[:, 1]) < tol, 0.0, np.where(np.abs(pts[:, 1] - Ly) < tol, Ly, pts[:, 1]))
    return pts


def triangulate_rect(Lx, Ly, fh, h0, loops=(), seed=0, maxit=300, fscale=1.2, dt=0.2):
    """Triangulate [0,Lx]x[0,Ly] with edge length ~ ``fh(points)`` (>= h0).  ``loops``: closed polylines whose vertices are
    fixed and whose edges appear in the result.  Returns ``(points [n,2], triangles [m,3] ccw, info)``, points
    numbered by position (x, then y)."""
    rng = np.random.default_rng(seed)
    fixed = [np.asarray(l, dtype=np.float64) for l in loops]
    corners = np.array([[0.0, 0.0], [Lx, 0.0], [Lx, Ly], [0.0, Ly]])
    pfix = np.concatenate(fixed + [corners], axis=0)
    nfix = pfix.shape[0]
    # thinned hexagonal lattice
    xs = np.arange(0.0, Lx + 0.5 * h0, h0)
    ys = np.arange(0.0, Ly + 0.5 * h0, h0 * _SQ3 / 2)
    X, Y = np.meshgrid(xs, ys, indexing="xy")
    X = X.copy(); X[1::2] += 0.5 * h0
    p = np.stack([X.ravel(), Y.ravel()], axis=1)
    p = p[(p[:, 0] <= Lx) & (p[:, 1] <= Ly)]
    keep = rng.random(p.shape[0]) < (h0 / fh(p)) ** 2
    p = p[keep]
    # no free point closer than 0.7 local sizes to a fixed one
    d2 = ((p[:, None, :] - pfix[None, :, :]) ** 2).sum(axis=2).min(axis=1)
    p = p[d2 > (0.7 * fh(p)) ** 2]
    p = np.concatenate([pfix, p], axis=0)
    lo = np.zeros(2); hi = np.array([Lx, Ly])
    pold = np.full_like(p, np.inf)
    tri = bars = None
    it = 0
    for it in range(maxit):
        if tri is None or np.sqrt(((p - pold) ** 2).sum(axis=1)).max() > 0.1 * h0:
            pold = p.copy()
            tri = _delaunay(p, h0)
            bars = _edges_of(tri)
        vec = p[bars[:, 0]] - p[bars[:, 1]]
        L = np.sqrt((vec ** 2).sum(axis=1))
        hb = fh(0.5 * (p[bars[:, 0]] + p[bars[:, 1]]))
        L0 = hb * fscale * np.sqrt((L ** 2).sum() / (hb ** 2).sum())
        F = np.maximum(L0 - L, 0.0)
        fv = (F / np.maximum(L, 1e-300))[:, None] * vec
        tot = np.zeros_like(p)
        np.add.at(tot, bars[:, 0], fv)
        np.add.at(tot, bars[:, 1], -fv)
        tot[:nfix] = 0.0
        step = dt * tot
        p = np.clip(p + step, lo, hi)                      # the domain is a rectangle: projection = clamping
        interior = (p[:, 0] > 0) & (p[:, 0] < Lx) & (p[:, 1] > 0) & (p[:, 1] < Ly)
        interior[:nfix] = False
        if interior.any() and np.sqrt((step[interior] ** 2).sum(axis=1)).max() < 2e-3 * h0:
            break
    # free points that ended up a hair inside the boundary go onto it
    free = p[nfix:]
    hloc = fh(free)
    for ax, L_ in ((0, Lx), (1, Ly)):
        free[:, ax] = np.where(free[:, ax] < 0.25 * hloc, 0.0, np.where(L_ - free[:, ax] < 0.25 * hloc, L_, free[:, ax]))
    # loop edges must be edges: empty their diametral circles of free points, then triangulate for good
    removed = 0
    if fixed:
        a = np.concatenate(fixed, axis=0)
        b = np.concatenate([np.roll(l, -1, axis=0) for l in fixed], axis=0)
        mid = 0.5 * (a + b)
        rad = 0.5 * np.sqrt(((a - b) ** 2).sum(axis=1)) * 1.02
        free = p[nfix:]
        d = np.sqrt(((free[:, None, :] - mid[None, :, :]) ** 2).sum(axis=2))
        bad = (d < rad[None, :]).any(axis=1)
        removed = int(bad.sum())
        p = np.concatenate([p[:nfix], free[~bad]], axis=0)
    tri = _delaunay(p, h0)
    if fixed:
        have = set(map(tuple, _edges_of(tri).tolist()))
        off = 0
        for l in fixed:
            n = l.shape[0]
            for i in range(n):
                e = (off + i, off + (i + 1) % n)
                if (min(e), max(e)) not in have:
                    raise RuntimeError("pad rim edge missing from the triangulation")
            off += n
    # number the points by position (x, then y).  The prism split downstream sends every quad diagonal from the lower-numbered
    # column up to the higher-numbered one: with this order all diagonals lean the same way, as in a Kuhn split.  Leaving the
    # rim vertices first made every element at a pad rim lean on the rim column, and nodal averages of the current on the pad
    # face over-weighted the rim (pad-current integral +8 % instead of +1 %).
    order = np.lexsort((p[:, 1], p[:, 0]))
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    p = p[order]
    tri = rank[tri]
    a, b, c = p[tri[:, 0]], p[tri[:, 1]], p[tri[:, 2]]
    la, lb, lc_ = (np.sqrt(((b - c) ** 2).sum(1)), np.sqrt(((a - c) ** 2).sum(1)), np.sqrt(((a - b) ** 2).sum(1)))
    q = 4 * _SQ3 * _tri_area(p, tri) / (la ** 2 + lb ** 2 + lc_ ** 2)   # 1 for an equilateral triangle
    info = dict(iterations=it + 1, points=int(p.shape[0]), triangles=int(tri.shape[0]), removed_near_rim=removed,
                min_quality=float(q.min()), mean_quality=float(q.mean()))
    return p, tri, info


# ---------------------------------------------------------------------------
# extrusion
# ---------------------------------------------------------------------------
def _prism_tets(tri, n2, k):
    """Three tets per prism between levels k and k+1 (node id = level * n2 + column).  Diagonals: lower column's bottom ->
    higher column's top."""
    t = tri.copy()
    # rotate each triangle so that its smallest column comes first (orientation kept)
    am = np.argmin(t, axis=1)
    idx = (am[:, None] + np.arange(3)[None, :]) % 3
    t = np.take_along_axis(t, idx, axis=1)
    v0, v1, v2 = t[:, 0] + k * n2, t[:, 1] + k * n2, t[:, 2] + k * n2
    w0, w1, w2 = v0 + n2, v1 + n2, v2 + n2
    lo12 = t[:, 1] < t[:, 2]
    ta = np.where(lo12[:, None], np.stack([v0, v1, v2, w2], 1), np.stack([v0, v1, v2, w1], 1))
    tb = np.where(lo12[:, None], np.stack([v0, v1, w2, w1], 1), np.stack([v0, v2, w2, w1], 1))
    tc = np.stack([v0, w1, w2, w0], 1)
    return np.concatenate([ta, tb, tc], axis=0)


def _wall_tris(edges, n2, k):
    """Two triangles per vertical quad over 2-D edges ``edges`` [m,2] between levels k and k+1, same diagonal rule."""
    e = np.sort(edges, axis=1)
    p, q = e[:, 0] + k * n2, e[:, 1] + k * n2
    return np.concatenate([np.stack([p, q, q + n2], 1), np.stack([p, q + n2, p + n2], 1)], axis=0)


def _boundary_edges(tri):
    e = np.concatenate([tri[:, [0, 1]], tri[:, [1, 2]], tri[:, [2, 0]]], axis=0)
    e = np.sort(e, axis=1)
    u, cnt = np.unique(e, axis=0, return_counts=True)
    return u[cnt == 1]


def graded_levels(z_top, z_bot, h_of_depth, depth0=0.0):
    """Levels from z_top down to z_bot whose spacing follows ``h_of_depth(depth0 + z_top - z)``, rescaled to end exactly
    at z_bot.  Returned ascending."""
    span = z_top - z_bot
    s = [0.0]
    while s[-1] < span - 1e-12:
        h = float(h_of_depth(depth0 + s[-1]))
        s.append(s[-1] + h)
    s = np.array(s)
    if len(s) > 2 and (s[-1] - span) > 0.5 * (s[-1] - s[-2]):
        s = s[:-1]                                       # the last step would be a sliver: absorb it
    s *= span / s[-1]
    z = z_top - s
    z[-1] = z_bot
    return z[::-1]


def layered_slab_graded(Lx=0.080, Ly=0.060, Lz=0.040, t_skin=0.0015, t_fat=0.005, t_contact=0.0005,
                        active_xy=(0.015, 0.045), return_xy=(0.065, 0.045), elec_r=0.010, shape="circle",
                        lc_elec=0.0015, lc_bulk=0.003, n_skin=None, n_fat=None, n_contact=1, contact_enabled=True,
                        interfaces_as_103=True, with_parents=True, bone=None, seed=0, plan=None, z_size_factor=1.0,
                        z_volume_law=False, z_size_factor_fat=None):
    """Layered slab with two contact pads (same geometry and tags as ``meshgen.layered_slab_mesh``,
    ``run_layered_sweep.py:142-181,206-227,296-308``) on a size-field-graded unstructured mesh.

    ``plan``: a ``(points, triangles, info)`` triple from an earlier call (``mesh.meta["plan"]``) to re-use the 2-D
    triangulation - it depends on the pad radius only, not on the layer thicknesses."""
    t_muscle = Lz - t_skin - t_fat
    if t_muscle <= 1e-4:
        raise ValueError(f"t_muscle = {t_muscle*1000:.2f} mm <= 0.1 mm - reduce t_fat + t_skin or increase Lz")
    centers = (tuple(active_xy), tuple(return_xy))

    def fh(p):
        return threshold(pad_distance_2d(p, centers, elec_r, shape), lc_elec, lc_bulk, elec_r, 6 * elec_r)
    if plan is None:
        loops = [pad_loop(c, elec_r, lc_elec, shape, Lx, Ly) for c in centers]
        plan = triangulate_rect(Lx, Ly, fh, lc_elec, loops, seed=seed)
    p2, tri, info = plan
    n2 = p2.shape[0]
    cen = p2[tri].mean(axis=1)

    def inside(c):
        if shape == "square":
            return (np.abs(cen[:, 0] - c[0]) < elec_r) & (np.abs(cen[:, 1] - c[1]) < elec_r)
        return np.hypot(cen[:, 0] - c[0], cen[:, 1] - c[1]) < elec_r
    m1, m2 = inside(centers[0]), inside(centers[1])
    if (m1 & m2).any():
        raise ValueError("electrode footprints overlap")
    area2 = _tri_area(p2, tri)
    # levels: muscle graded away from the pads, fat and skin uniform
    z0_fat, z0_skin = t_muscle, t_muscle + t_fat
    zf_fat = z_size_factor if z_size_factor_fat is None else z_size_factor_fat
    nf = n_fat if n_fat else max(2, int(round(t_fat / (lc_elec * zf_fat))))
    ns = n_skin if n_skin else max(1, int(round(t_skin / lc_elec)))
    nc = n_contact if contact_enabled else 0
    depth0 = t_skin + t_fat + (t_contact if nc else 0.0)
    def h_muscle(d):
        lc = threshold(d, lc_elec, lc_bulk, elec_r, 6 * elec_r)
        if z_volume_law:
            # the extrusion cannot coarsen the triangles under a pad with depth (they keep lc_elec where the reference's
            # 3-D size field has grown to lc(d)): the level spacing makes up for it, so that a prism under the pad has the
            # volume lc(d)^3 asks for - bounded by twice the bulk size
            lc = min(lc * (lc / lc_elec) ** 2, 2.0 * lc_bulk)
        return z_size_factor * lc
    zm = graded_levels(z0_fat, 0.0, h_muscle, depth0)
    zl = [zm, np.linspace(z0_fat, z0_skin, nf + 1)[1:], np.linspace(z0_skin, Lz, ns + 1)[1:]]
    if nc:
        zl.append(np.linspace(Lz, Lz + t_contact, nc + 1)[1:])
    zs = np.concatenate(zl)
    n_m = len(zm) - 1
    k_fat, k_skin, k_top = n_m, n_m + nf, n_m + nf + ns
    zs[k_fat] = z0_fat; zs[k_skin] = z0_skin; zs[k_top] = Lz
    nlev = len(zs)
    nodes = np.empty((nlev * n2, 3))
    nodes[:, 0] = np.tile(p2[:, 0], nlev)
    nodes[:, 1] = np.tile(p2[:, 1], nlev)
    nodes[:, 2] = np.repeat(zs, n2)
    tets, region = [], []
    bone_meta = None
    inb2 = None
    if bone is not None:
        # bone block: the triangles whose centroid lies in the x-y extent, between the nearest muscle levels
        ka = int(np.argmin(np.abs(zs[:k_fat + 1] - bone["z"][0]))); kb = int(np.argmin(np.abs(zs[:k_fat + 1] - bone["z"][1])))
        if kb <= ka:
            raise ValueError("bone block is thinner than one level of this mesh")
        inb2 = ((cen[:, 0] > bone["x"][0]) & (cen[:, 0] < bone["x"][1]) & (cen[:, 1] > bone["y"][0]) & (cen[:, 1] < bone["y"][1]))
        bone_meta = dict(x=tuple(bone["x"]), y=tuple(bone["y"]), z=(float(zs[ka]), float(zs[kb])), levels=(ka, kb))
    for k in range(nlev - 1):
        if k < k_top:
            tt = _prism_tets(tri, n2, k)
            body = np.full(tri.shape[0], 1 if k < k_fat else 2 if k < k_skin else 3, dtype=np.int32)
            if inb2 is not None and bone_meta["levels"][0] <= k < bone_meta["levels"][1]:
                body = np.where(inb2, 6, body).astype(np.int32)
            tets.append(tt); region.append(np.tile(body, 3))
        else:
            for msk, b in ((m1, 4), (m2, 5)):
                tt = _prism_tets(tri[msk], n2, k)
                tets.append(tt); region.append(np.full(tt.shape[0], b, dtype=np.int32))
    tets = np.concatenate(tets, axis=0)
    region = np.concatenate(region)
    # boundary triangles
    def level_tris(k, msk=None):
        t = tri if msk is None else tri[msk]
        return t + k * n2
    k_elec = nlev - 1 if nc else k_top
    t101, t102 = level_tris(k_elec, m1), level_tris(k_elec, m2)
    others = [level_tris(0), level_tris(k_top, ~(m1 | m2))]
    if nc:
        for msk in (m1, m2):
            rim = _boundary_edges(tri[msk])
            for k in range(k_top, nlev - 1):
                others.append(_wall_tris(rim, n2, k))
        if interfaces_as_103:
            others.append(level_tris(k_top, m1 | m2))
    if interfaces_as_103:
        others.append(level_tris(k_fat)); others.append(level_tris(k_skin))
    outer = _boundary_edges(tri)
    for k in range(k_top):
        others.append(_wall_tris(outer, n2, k))
    t103 = np.concatenate(others, axis=0)
    tris = np.concatenate([t101, t102, t103], axis=0)
    bcid = np.concatenate([np.full(len(t101), 101), np.full(len(t102), 102), np.full(len(t103), 103)]).astype(np.int32)
    used = np.zeros(nodes.shape[0], dtype=bool)
    used[tets.ravel()] = True
    new_id = np.cumsum(used) - 1
    nodes = np.ascontiguousarray(nodes[used])
    tets = np.ascontiguousarray(new_id[tets].astype(np.int32))
    tris = np.ascontiguousarray(new_id[tris].astype(np.int32))
    orient_positive(nodes, tets)
    if not (tet_volumes(nodes, tets) > 0).all():
        raise RuntimeError("degenerate element in the extruded mesh")
    z_elec_top = Lz + (t_contact if nc else 0.0)
    mesh = TetMesh(nodes, tets, region.astype(np.int32), tris, bcid,
                   meta=dict(kind="layered_slab_graded", Lx=Lx, Ly=Ly, Lz=Lz, t_skin=t_skin, t_fat=t_fat,
                             t_contact=t_contact if nc else 0.0, elec_r=elec_r, shape=shape, active_xy=centers[0],
                             return_xy=centers[1], z_elec_top=z_elec_top, contact_enabled=bool(nc), bone=bone_meta,
                             area_active=float(area2[m1].sum()), area_return=float(area2[m2].sum()),
                             levels=zs.copy(), columns=n2, plan=plan, triangulation=info))
    if with_parents:
        mesh.tri_parent = find_tri_parents(mesh.tets, mesh.tris)
    return mesh


def electrode_box_graded(Lx, Ly, Lz, e1_xy, e2_xy, r, shape="circle", lc_elec=None, lc_bulk=None, dist_max_factor=7.0,
                         z_size_factor=1.0, with_parents=True, seed=0):
    """Homogeneous box with two electrode patches on its top face (geometry and tags of
    ``step02_electrodes/run_sweep.py:39-52,63-103``; sizes ``:108-119``: ``r/3.5`` up to distance ``r`` from the patches,
    growing to ``min(4 r, 12 mm)`` at ``7 r``) on the size-field mesh: graded 2-D triangulation with the patch outlines as
    polygons, extruded downwards with level spacing following the same law in depth."""
    lc_elec = lc_elec if lc_elec is not None else r / 3.5
    lc_bulk = lc_bulk if lc_bulk is not None else min(4 * r, 0.012)
    centers = (tuple(e1_xy), tuple(e2_xy))

    def fh(p):
        return threshold(pad_distance_2d(p, centers, r, shape), lc_elec, lc_bulk, r, dist_max_factor * r)
    loops = [pad_loop(c, r, lc_elec, shape, Lx, Ly) for c in centers]
    p2, tri, info = triangulate_rect(Lx, Ly, fh, lc_elec, loops, seed=seed)
    n2 = p2.shape[0]
    cen = p2[tri].mean(axis=1)

    def inside(c):
        if shape == "square":
            return (np.abs(cen[:, 0] - c[0]) < r) & (np.abs(cen[:, 1] - c[1]) < r)
        return np.hypot(cen[:, 0] - c[0], cen[:, 1] - c[1]) < r
    m1, m2 = inside(centers[0]), inside(centers[1])
    zs = graded_levels(Lz, 0.0, lambda d: z_size_factor * threshold(d, lc_elec, lc_bulk, r, dist_max_factor * r))
    nlev = len(zs)
    nodes = np.empty((nlev * n2, 3))
    nodes[:, 0] = np.tile(p2[:, 0], nlev); nodes[:, 1] = np.tile(p2[:, 1], nlev); nodes[:, 2] = np.repeat(zs, n2)
    tets = np.concatenate([_prism_tets(tri, n2, k) for k in range(nlev - 1)], axis=0).astype(np.int32)
    orient_positive(nodes, tets)
    top = (nlev - 1) * n2
    outer = _boundary_edges(tri)
    t101, t102 = tri[m1] + top, tri[m2] + top
    t103 = np.concatenate([tri[~(m1 | m2)] + top, tri] + [_wall_tris(outer, n2, k) for k in range(nlev - 1)], axis=0)
    tris = np.concatenate([t101, t102, t103], axis=0).astype(np.int32)
    bcid = np.concatenate([np.full(len(t101), 101), np.full(len(t102), 102), np.full(len(t103), 103)]).astype(np.int32)
    area2 = _tri_area(p2, tri)
    mesh = TetMesh(nodes, np.ascontiguousarray(tets), np.ones(tets.shape[0], dtype=np.int32), np.ascontiguousarray(tris), bcid,
                   meta=dict(kind="electrode_box_graded", Lx=Lx, Ly=Ly, Lz=Lz, r=r, shape=shape, levels=zs.copy(), columns=n2,
                             area_active=float(area2[m1].sum()), area_return=float(area2[m2].sum()), triangulation=info))
    if with_parents:
        mesh.tri_parent = find_tri_parents(mesh.tets, mesh.tris)
    return mesh
