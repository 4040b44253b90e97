"""CPU ORACLE (test infrastructure only) — numpy restatement of the reference's
metric extraction (layer L5), which it implements with pyvista/VTK filters on
``results/case_t0001.vtu``.  pyvista is not available here, so the VTK filter
semantics are restated explicitly:

* ``point_data_to_cell_data``: cell value = arithmetic mean of the cell's nodes.
* ``compute_derivative`` on a CELL-data scalar: vtkGradientFilter first maps the
  cell data back to points (vtkCellDataToPointData: unweighted mean over all
  cells that use the point), then evaluates the linear cell's derivative at the
  cell centre from those point values (tet: sum_i v_i grad N_i; triangle:
  in-plane gradient).
* ``cell_centers`` = mean of node coordinates; ``compute_cell_sizes`` Area of a
  triangle = |cross|/2.

Cells are the tets followed by the boundary triangles (VTK types 10 then 5).

PARITY STATUS: unpinned against pyvista itself (not installed); pinned only by
the definitions above and by hand-computed cases in tests/.
"""
from __future__ import annotations

import numpy as np

from .fem_oracle import tet_geometry


# -- A8: step01 / step02 ----------------------------------------------------------
def step01_metrics(pts, phi, J, sigma=0.2, v_top=1.0, v_bot=0.0):
    """``step01_box/test_step01_baseline.py:59-104``."""
    Jmag = np.linalg.norm(J, axis=1)
    Lx, Ly, Lz = pts[:, 0].max(), pts[:, 1].max(), pts[:, 2].max()
    J_an = sigma * (v_top - v_bot) / Lz
    mean_J = Jmag.mean()
    std_J = Jmag.std(ddof=1)
    r_xy = np.hypot(pts[:, 0] - Lx / 2, pts[:, 1] - Ly / 2)
    col = r_xy < Lx * 0.08
    z_c, phi_c = pts[col, 2], phi[col]
    coeffs = np.polyfit(z_c, phi_c, 1)
    fit = np.polyval(coeffs, z_c)
    r2 = 1.0 - np.sum((phi_c - fit) ** 2) / np.sum((phi_c - phi_c.mean()) ** 2)
    tol_z = Lz * 1e-3
    ft = np.abs(J[pts[:, 2] > Lz - tol_z, 2]).mean()
    fb = np.abs(J[pts[:, 2] < tol_z, 2]).mean()
    return dict(Lz=Lz, J_an=J_an, mean_J=mean_J, std_J=std_J, cv_J=std_J / mean_J,
                rel_J=abs(mean_J - J_an) / J_an, r2=r2, slope=coeffs[0], flux_top=ft, flux_bot=fb,
                flux_err=abs(ft - fb) / max(ft, fb), phi_min=phi.min(), phi_max=phi.max())


def top_face_J(pts, J, Lz):
    """``step02_electrodes/run_sweep.py:286-295,331-333``: nodes with z > 0.99 Lz;
    peak = max |J|, mean = unweighted node mean over the whole top face."""
    Jmag = np.linalg.norm(J, axis=1)
    m = pts[:, 2] > Lz * 0.99
    return float(Jmag[m].max()), float(Jmag[m].mean()), int(m.sum())


# -- A9: electrode currents ---------------------------------------------------------
def injected_current(pts, tris, J, e1_pos, e2_pos, elec_r, z_e1_top, z_e2_top, shape="circle"):
    """``step03_ankle_layers/run_layered_sweep.py:704-761``: on the boundary triangle
    cells, J_cell = mean of nodal J, I = sum J_z * area over cells whose centroid is on
    the electrode top (z > z_top - max(5e-3 z_top, 1e-5)) and within 1.2 r of the pad centre."""
    tol = 0.2
    Jc = J[tris].mean(axis=1)
    cen = pts[tris].mean(axis=1)
    p = pts[tris]
    area = 0.5 * np.linalg.norm(np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]), axis=1)

    def mask(pos, z_top):
        tol_z = max(z_top * 5e-3, 1e-5)
        top = cen[:, 2] > z_top - tol_z
        dx, dy = cen[:, 0] - pos[0], cen[:, 1] - pos[1]
        if shape == "square":
            return top & (np.abs(dx) < elec_r * (1 + tol)) & (np.abs(dy) < elec_r * (1 + tol))
        return top & (np.sqrt(dx ** 2 + dy ** 2) < elec_r * (1 + tol))
    ma, mr = mask(e1_pos, z_e1_top), mask(e2_pos, z_e2_top)
    if not ma.any() or not mr.any():
        return (np.nan,) * 5
    Ia = float(np.sum(Jc[ma, 2] * area[ma]))
    Ir = float(np.sum(Jc[mr, 2] * area[mr]))
    denom = max(abs(Ia), abs(Ir))
    ferr = float(abs(Ia + Ir) / denom) if denom > 0 else np.nan
    return abs(Ia), abs(Ir), ferr, Ia, Ir


# -- A10: ROI ---------------------------------------------------------------------------
def cell_fields(pts, tets, tris, phi, J):
    """Cell-centre |J|, |E| and centroids for all cells (tets then tris) with the VTK
    semantics described in the module docstring."""
    nn = pts.shape[0]
    Jc = np.concatenate([J[tets].mean(axis=1), J[tris].mean(axis=1)], axis=0)
    phic_t = phi[tets].mean(axis=1)
    phic_b = phi[tris].mean(axis=1)
    acc = np.zeros(nn)
    cnt = np.zeros(nn)
    for a in range(4):          # bincount adds in input order, like np.add.at, but vectorised
        acc += np.bincount(tets[:, a], weights=phic_t, minlength=nn)
        cnt += np.bincount(tets[:, a], minlength=nn)
    for a in range(3):
        acc += np.bincount(tris[:, a], weights=phic_b, minlength=nn)
        cnt += np.bincount(tris[:, a], minlength=nn)
    phis = acc / np.maximum(cnt, 1.0)                   # smoothed point values
    _, g = tet_geometry(pts, tets)
    grad_t = np.einsum("ei,eik->ek", phis[tets], g)
    p = pts[tris]
    e1 = p[:, 1] - p[:, 0]
    e2 = p[:, 2] - p[:, 0]
    n = np.cross(e1, e2)
    n2 = np.einsum("ij,ij->i", n, n)
    n2 = np.where(n2 > 0, n2, 1.0)
    v = phis[tris]
    # in-plane gradient of the linear interpolant on a triangle
    grad_b = ((v[:, 1] - v[:, 0])[:, None] * np.cross(e2, n) + (v[:, 2] - v[:, 0])[:, None] * np.cross(n, e1)) / n2[:, None]
    E = -np.concatenate([grad_t, grad_b], axis=0)
    cen = np.concatenate([pts[tets].mean(axis=1), pts[tris].mean(axis=1)], axis=0)
    return np.linalg.norm(Jc, axis=1), np.linalg.norm(E, axis=1), cen


def eval_roi(pts, tets, tris, phi, J, roi_cen, roi_radius_init, min_cells=4):
    """``run_layered_sweep.py:765-822``: mean |J_c| and |E_c| over cells whose centre lies in
    the ROI sphere; radius expands x1.5/2/3 until >= min_cells cells."""
    Jm, Em, cen = cell_fields(pts, tets, tris, phi, J)
    dist = np.linalg.norm(cen - np.asarray(roi_cen), axis=1)
    used = roi_radius_init
    mask = None
    for mult in (1.0, 1.5, 2.0, 3.0):
        r_test = roi_radius_init * mult
        mask = dist < r_test
        if int(mask.sum()) >= min_cells:
            used = r_test
            break
    else:
        used = roi_radius_init * 3.0
        mask = dist < used
    n = int(mask.sum())
    if n == 0:
        return np.nan, np.nan, 0, used, cen, mask
    return float(Jm[mask].mean()), float(Em[mask].mean()), n, used, cen, mask


# -- A11: the summary rows -----------------------------------------------------------------
def _r(val, n):
    v = float(val)
    return round(v, n) if np.isfinite(v) else v


def skin_peaks(pts, J, z0_skin, t_skin, e1_pos, e2_pos, elec_r, shape):
    """``run_layered_sweep.py:849-871``."""
    Jmag = np.linalg.norm(J, axis=1)
    skin = pts[:, 2] > z0_skin + t_skin * 0.80
    if not skin.any():
        return np.nan, np.nan
    peak_with = float(Jmag[skin].max())
    xp, yp, Jm = pts[skin, 0], pts[skin, 1], Jmag[skin]

    def inside(xc, yc):
        if shape == "circle":
            return np.sqrt((xp - xc) ** 2 + (yp - yc) ** 2) < elec_r
        return (np.abs(xp - xc) < elec_r) & (np.abs(yp - yc) < elec_r)
    out = ~(inside(e1_pos[0], e1_pos[1]) | inside(e2_pos[0], e2_pos[1]))
    peak_no = float(Jm[out].max()) if out.any() else peak_with
    return peak_with, peak_no


def compliance_voltage(pts, phi, e1_pos, e2_pos, elec_r, z_e1_top, z_e2_top, shape):
    """``run_layered_sweep.py:899-920``."""
    def mask(pos, z_et):
        tol_z = max(z_et * 5e-3, 1e-5)
        m = pts[:, 2] > z_et - tol_z
        if shape == "circle":
            m &= np.sqrt((pts[:, 0] - pos[0]) ** 2 + (pts[:, 1] - pos[1]) ** 2) < elec_r * 1.5
        else:
            m &= (np.abs(pts[:, 0] - pos[0]) < elec_r * 1.5) & (np.abs(pts[:, 1] - pos[1]) < elec_r * 1.5)
        return m
    ma, mr = mask(e1_pos, z_e1_top), mask(e2_pos, z_e2_top)
    if not ma.any():
        return np.nan
    return float(phi[ma].mean()) - (float(phi[mr].mean()) if mr.any() else 0.0)


def layered_row(pts, tets, tris, phi, J, p, t_fat, elec_r, e1_pos, e2_pos, body_info,
                sigma_skin_used=None, jn_used=None, elec_area_mesh=None, return_area_mesh=None,
                e1_id=None, e2_id=None):
    """The 36-column step03 row: ``run_layered_sweep.py:826-1030``."""
    ls = p["layers"]
    st = p.get("stim", p.get("control", {}))
    z_skin_top = body_info["z_skin_top"]
    z1 = body_info.get("z_e1_elec_top", body_info["z_elec_top"])
    z2 = body_info.get("z_e2_elec_top", body_info["z_elec_top"])
    shape = body_info.get("elec_shape", "circle")
    z0_skin = z_skin_top - ls["t_skin"]
    peak_with, peak_no = skin_peaks(pts, J, z0_skin, ls["t_skin"], e1_pos, e2_pos, elec_r, shape)
    Ia, Ir, ferr, Ias, Irs = injected_current(pts, tris, J, e1_pos, e2_pos, elec_r, z1, z2, shape)
    mode = st.get("control_mode", "voltage")
    comp = np.nan
    exceeded = False
    if mode == "current":
        comp = compliance_voltage(pts, phi, e1_pos, e2_pos, elec_r, z1, z2, shape)
        if np.isfinite(comp):
            exceeded = bool(comp > st.get("compliance_voltage_V", 100.0))
    rc = p["roi"]
    z_nerve = z_skin_top - rc["z_target"]
    roi_cen = np.array([e1_pos[0], e1_pos[1], z_nerve])
    mJ, mE, ncell, r_used, cen, _ = eval_roi(pts, tets, tris, phi, J, roi_cen, rc["roi_radius"])
    t_sk = ls["t_skin"]
    z_fat_bot = z_skin_top - t_sk - t_fat
    z_fat_top = z_skin_top - t_sk
    dist_all = np.linalg.norm(cen - roi_cen, axis=1)
    m_all = dist_all < r_used
    if m_all.any():
        z_roi = cen[m_all, 2]
        n_roi = m_all.sum()
        f_skin = float((z_roi > z_fat_top).sum()) / n_roi
        f_fat = float(((z_roi > z_fat_bot) & (z_roi <= z_fat_top)).sum()) / n_roi
        f_mus = float((z_roi <= z_fat_bot).sum()) / n_roi
    else:
        f_skin = f_fat = f_mus = np.nan
    area = np.pi * elec_r ** 2 if shape == "circle" else (2 * elec_r) ** 2
    eff = float(mE) / peak_no if (np.isfinite(mE) and peak_no > 0) else np.nan
    I_ref = Ia if np.isfinite(Ia) and Ia > 0 else np.nan

    def norm(v):
        v = float(v)
        return v / I_ref if np.isfinite(v) and np.isfinite(I_ref) else np.nan
    roi_layer = "skin" if z_nerve > z_skin_top - t_sk else "fat" if z_nerve > z_fat_bot else "muscle"
    sig = sigma_skin_used if sigma_skin_used is not None else p["conductivities"]["sigma_skin"]
    return {
        "t_fat_mm": _r(t_fat * 1000, 2), "elec_r_mm": _r(elec_r * 1000, 2), "elec_area_cm2": _r(area * 1e4, 4),
        "elec_area_mesh_cm2": _r(elec_area_mesh * 1e4, 4) if elec_area_mesh else None,
        "return_area_mesh_cm2": _r(return_area_mesh * 1e4, 4) if return_area_mesh else None,
        "elec_shape": shape, "contact_enabled": body_info.get("contact_enabled", False), "sigma_skin": sig,
        "control_mode": mode, "jn_used": _r(jn_used, 4) if jn_used is not None else None,
        "peak_J_skin_with_elec": _r(peak_with, 6), "peak_J_skin_no_elec": _r(peak_no, 6),
        "roi_mean_J": _r(mJ, 6), "roi_mean_E": _r(mE, 4), "efficiency": _r(eff, 6),
        "compliance_V": _r(comp, 3), "exceeded_compliance": exceeded,
        "total_current_A": _r(Ia, 8), "I_active_signed_A": _r(Ias, 8), "I_return_A": _r(Ir, 8),
        "I_return_signed_A": _r(Irs, 8), "peak_J_skin_per_A": _r(norm(peak_no), 4),
        "roi_mean_J_per_A": _r(norm(mJ), 4), "roi_mean_E_per_A": _r(norm(mE), 4), "efficiency_per_A": _r(eff, 6),
        "flux_err": _r(ferr, 6), "roi_layer": roi_layer, "roi_n_cells": ncell,
        "roi_radius_used_mm": _r(r_used * 1000, 2), "roi_center_z_mm": _r(z_nerve * 1000, 3),
        "dist_fat_muscle_mm": _r(abs(z_nerve - z_fat_bot) * 1000.0, 3), "roi_frac_muscle": _r(f_mus, 4),
        "roi_frac_fat": _r(f_fat, 4), "roi_frac_skin": _r(f_skin, 4),
        "active_boundary_id_used": e1_id, "return_boundary_id_used": e2_id,
    }


def pressure_row(pts, tets, tris, phi, J, p, sigma_contact, label, e1_pos, e2_pos, body_info, jn_used):
    """The 24-column step04 row: ``step04_pressure/run_pressure_sweep.py:528-660``."""
    ls, st = p["layers"], p.get("stim", p.get("control", {}))
    pl = p.get("placement", p.get("electrodes", {}))
    elec_r = float(pl["electrode_r_mm"]) * 1e-3
    shape = body_info["elec_shape"]
    z_skin_top = body_info["z_skin_top"]
    z1, z2 = body_info["z_e1_elec_top"], body_info["z_e2_elec_top"]
    t_skin = ls["t_skin"]
    peak_with, peak_no = skin_peaks(pts, J, z_skin_top - t_skin, t_skin, e1_pos, e2_pos, elec_r, "circle")
    Ia, Ir, ferr, Ias, Irs = injected_current(pts, tris, J, e1_pos, e2_pos, elec_r, z1, z2, shape)
    # compliance uses the active electrode's z tolerance for both masks (``:583-590``) and always circles
    tol_z = max(z1 * 5e-3, 1e-5)
    act = (pts[:, 2] > z1 - tol_z) & (np.sqrt((pts[:, 0] - e1_pos[0]) ** 2 + (pts[:, 1] - e1_pos[1]) ** 2) < elec_r * 1.5)
    ret = (pts[:, 2] > z2 - tol_z) & (np.sqrt((pts[:, 0] - e2_pos[0]) ** 2 + (pts[:, 1] - e2_pos[1]) ** 2) < elec_r * 1.5)
    comp = np.nan
    if act.any():
        comp = float(phi[act].mean()) - (float(phi[ret].mean()) if ret.any() else 0.0)
    exceeded = bool(np.isfinite(comp) and comp > st.get("compliance_voltage_V", 200.0))
    Z = float(comp / Ia) if (np.isfinite(comp) and np.isfinite(Ia) and Ia > 0) else np.nan
    rc = p["roi"]
    roi_cen = np.array([e1_pos[0], e1_pos[1], z_skin_top - rc["z_target"]])
    mJ, mE, ncell, r_used, _, _ = eval_roi(pts, tets, tris, phi, J, roi_cen, rc["roi_radius"])
    pw_us = st.get("pulse_width_us", 200.0)
    charge = float(peak_with * pw_us * 1e-6 * 0.1) if np.isfinite(peak_with) else np.nan
    limit = p.get("safety", {}).get("charge_density_limit_mC_cm2", 1.0)
    eff = float(mE) / peak_no if (np.isfinite(mE) and peak_no > 0) else np.nan
    return {
        "pressure_label": label, "sigma_contact_Spm": sigma_contact, "elec_r_mm": float(pl["electrode_r_mm"]),
        "t_fat_mm": ls["t_fat"] * 1000, "compliance_V": _r(comp, 3), "contact_impedance_ohm": _r(Z, 1),
        "exceeded_compliance": exceeded, "I_active_A": _r(Ia, 8), "I_return_A": _r(Ir, 8),
        "I_active_signed_A": _r(Ias, 8), "I_return_signed_A": _r(Irs, 8), "flux_err": _r(ferr, 6),
        "jn_used_A_m2": _r(jn_used, 6), "peak_J_skin_with_elec": _r(peak_with, 4),
        "peak_J_skin_no_elec": _r(peak_no, 4), "charge_density_mC_cm2": _r(charge, 6),
        "exceeds_charge_limit": bool(np.isfinite(charge) and charge > limit), "roi_mean_J": _r(mJ, 6),
        "roi_mean_E": _r(mE, 4), "efficiency": _r(eff, 6), "roi_n_cells": ncell,
        "roi_radius_used_mm": _r(r_used * 1000, 2), "pulse_width_us": pw_us,
        "frequency_Hz": st.get("frequency_Hz", 10.0),
    }


# -- A12: boundary-id detection (pure-Python restatement, small meshes only) ----------------------
def detect_elec_bc_ids_loops(nodes, tris, bcid, e1_pos, e2_pos, z_e1_top, z_e2_top):
    """``run_layered_sweep.py:366-455`` with plain loops (checker for the vectorised host code)."""
    z_floor = min(z_e1_top, z_e2_top) - 5e-3
    cen, zc, elems = {}, {}, {}
    for t, b in zip(tris, bcid):
        c = nodes[t]
        if c[:, 2].max() < z_floor:
            continue
        cen.setdefault(int(b), []).append(c[:, :2].mean(axis=0))
        zc.setdefault(int(b), []).append(float(c[:, 2].mean()))
        elems.setdefault(int(b), []).append(t)
    mean_xy = {b: np.mean(v, axis=0) for b, v in cen.items()}
    mean_z = {b: float(np.mean(v)) for b, v in zc.items()}
    if len(mean_xy) < 2:
        raise RuntimeError(f"Expected >=2 top-face BCs, found: {list(mean_xy)}")

    def find(pos, z_top, exclude=None):
        tol = max(z_top * 2e-2, 5e-4)
        cand = {b: c for b, c in mean_xy.items() if b != exclude and abs(mean_z.get(b, 0) - z_top) < tol}
        if not cand:
            cand = {b: c for b, c in mean_xy.items() if b != exclude}
        return min(cand, key=lambda b: np.linalg.norm(cand[b] - np.asarray(pos[:2])))
    e1 = find(e1_pos, z_e1_top)
    e2 = find(e2_pos, z_e2_top, exclude=e1)

    def area(b):
        a = 0.0
        for t in elems.get(b, []):
            v0, v1, v2 = nodes[t]
            a += 0.5 * float(np.linalg.norm(np.cross(v1 - v0, v2 - v0)))
        return a
    return e1, e2, area(e1), area(e2)
