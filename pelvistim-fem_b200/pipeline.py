"""Host-side pipeline shared by the four drivers: the pieces of the reference scripts that sit
either side of the ``ElmerSolver`` call, with the solve itself and every metric reduction done on
the GPU through ``engine`` (libptfem.so).

Mirrors, function by function (same names, argument meaning and error behaviour):

* ``detect_elec_bc_ids``   — ``step03_ankle_layers/run_layered_sweep.py:366-455``
* ``save_bc_debug_report`` — ``run_layered_sweep.py:647-700``
* ``run_elmer_solver``     — the ``ElmerSolver case.sif`` subprocess (``:1099``): reads ``case.sif`` +
  ``elmer_mesh/`` of a case directory, solves, writes ``results/case_t0001.vtu``
* ``extract_layered`` / ``extract_pressure`` / ``extract_top_J`` / ``step01_metrics`` — the metric
  extraction of ``run_layered_sweep.py:826-1030``, ``run_pressure_sweep.py:528-660``,
  ``run_sweep.py:286-295`` and ``test_step01_baseline.py:59-104``; pyvista filters are replaced by the
  K12 reduction kernels (same cell/point semantics, see ``oracle/metrics_oracle.py``).
"""
from __future__ import annotations

import math
from pathlib import Path

import numpy as np

from . import elmer_io, sif, vtu
from .engine import Context, DeviceMesh

NAN = float("nan")


# ---------------------------------------------------------------------------
# pre-solve: boundary ids + electrode mesh areas
# ---------------------------------------------------------------------------
def detect_elec_bc_ids(mesh_or_dir, e1_pos3d, e2_pos3d, z_e1_top, z_e2_top):
    """Pick the boundary ids of the active / return electrode patches geometrically and measure
    their mesh areas.  Returns ``(e1_id, e2_id, A_active, A_return)``."""
    mesh = mesh_or_dir if hasattr(mesh_or_dir, "tris") else elmer_io.read_elmer_mesh(mesh_or_dir)
    p = mesh.nodes[mesh.tris]                          # [nb,3,3]
    z_floor = min(z_e1_top, z_e2_top) - 5e-3
    near = p[:, :, 2].max(axis=1) >= z_floor
    if not near.any():
        raise RuntimeError("Expected >=2 top-face BCs, found: []")
    ids = np.unique(mesh.bcid[near])
    if ids.size < 2:
        raise RuntimeError(f"Expected >=2 top-face BCs, found: {ids.tolist()}")
    cen_xy = p[:, :, :2].mean(axis=1)
    cen_z = p[:, :, 2].mean(axis=1)
    mean_xy, mean_z = {}, {}
    for b in ids.tolist():
        m = near & (mesh.bcid == b)
        mean_xy[b] = cen_xy[m].mean(axis=0)
        mean_z[b] = float(cen_z[m].mean())

    def find(pos, z_top, exclude=None):
        tol = max(z_top * 2e-2, 5e-4)
        cand = [b for b in mean_xy if b != exclude and abs(mean_z[b] - z_top) < tol]
        if not cand:
            cand = [b for b in mean_xy if b != exclude]
        return min(cand, key=lambda b: float(np.linalg.norm(mean_xy[b] - np.asarray(pos[:2], dtype=float))))
    e1 = find(e1_pos3d, z_e1_top)
    e2 = find(e2_pos3d, z_e2_top, exclude=e1)
    area = 0.5 * np.linalg.norm(np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]), axis=1)

    def total(b):
        return float(area[near & (mesh.bcid == b)].sum())
    return int(e1), int(e2), total(e1), total(e2)


def classify_flat_boundaries(mesh):
    """``step01_box/setup_case.py:107-118``: boundary ids whose elements all lie on the global zmax
    (top) / zmin (bottom) plane.  Returns ``(top_ids, bottom_ids)``."""
    z = mesh.nodes[:, 2]
    zmin, zmax = float(z.min()), float(z.max())
    tol = 1e-6 * max(zmax - zmin, 1e-30)
    tz = mesh.nodes[mesh.tris][:, :, 2]
    top, bot = [], []
    for b in np.unique(mesh.bcid).tolist():
        zz = tz[mesh.bcid == b]
        if np.all(np.abs(zz - zmax) <= tol):
            top.append(int(b))
        elif np.all(np.abs(zz - zmin) <= tol):
            bot.append(int(b))
    return top, bot


def save_bc_debug_report(run_dir, label, e1_id, e2_id, A_active_mesh, A_return_mesh, jn_used, p, body_info):
    st = p.get("stim", p.get("control", {}))
    mode = st.get("control_mode", "voltage")
    I_mA = st.get("injected_current_mA", 5.0)
    I_A = I_mA * 1e-3
    L = [f"BC DEBUG REPORT — {label}", "=" * 60, f"  control_mode     : {mode}",
         f"  injected_current : {I_mA} mA  ({I_A:.4e} A)", "",
         f"  Elmer boundary ID — active  : {e1_id}", f"  Elmer boundary ID — return  : {e2_id}", "",
         f"  Mesh area — active electrode : {A_active_mesh*1e4:.4f} cm²",
         f"  Mesh area — return electrode : {A_return_mesh*1e4:.4f} cm²"]
    if mode == "current" and jn_used is not None:
        expected = jn_used * A_active_mesh
        L += ["", f"  Current density applied (Jn) : {jn_used:.6e} A/m²",
              f"  Expected current (Jn * A)    : {expected*1e3:.4f} mA",
              f"  Target current               : {I_mA:.4f} mA",
              f"  Pre-solve area error         : {abs(expected - I_A)/I_A*100:.2f}%", "",
              "  Elmer BC keyword used: 'Current Density = Jn'",
              "  Elmer interprets this as uniform normal J (A/m²) Neumann BC.",
              "  n_outward at top face = +z; current INTO tissue has J_z < 0.",
              "  This BC applies ONLY to the active electrode surface.",
              "  Return electrode is Dirichlet: Potential = 0."]
    zs = body_info["z_skin_top"]
    L += ["", f"  contact_enabled  : {body_info.get('contact_enabled', False)}",
          f"  z_skin_top (nom) : {zs*1000:.2f} mm",
          f"  z_e1_skin        : {body_info.get('z_e1_skin', zs)*1000:.2f} mm  (active electrode skin surface)",
          f"  z_e2_skin        : {body_info.get('z_e2_skin', zs)*1000:.2f} mm  (return electrode skin surface)",
          f"  z_e1_elec_top    : {body_info.get('z_e1_elec_top', body_info['z_elec_top'])*1000:.2f} mm",
          f"  z_e2_elec_top    : {body_info.get('z_e2_elec_top', body_info['z_elec_top'])*1000:.2f} mm"]
    out = Path(run_dir) / "bc_debug_report.txt"
    out.write_text("\n".join(L) + "\n")
    return out


# ---------------------------------------------------------------------------
# the solve step (what `ElmerSolver case.sif` does)
# ---------------------------------------------------------------------------
# Nodal "volume current" recovery used by the drivers and the ElmerSolver shim.  The reference's tables pin this
# choice: on the rim-fitted built-in meshes the pad current integrated from the nodal field is 5.58 / 5.30 / 5.20 mA
# (r = 5 / 10 / 15 mm) with the volume-weighted nodal average ("lumped"), 6.25 / 5.65 / 5.39 mA with the consistent-mass
# L2 projection, against 5.51 / 5.27 / 5.14 mA in step03_ankle_layers/results/summary.csv - Elmer's field behaves like
# the former.  ``solver: {current_recovery: l2 | lumped | average}`` in params.yaml overrides it.
DEFAULT_RECOVER = "lumped"


class SolvedCase:
    """Device-resident result of one case: mesh, solution and recovered current stay on the GPU
    for the metric kernels; ``phi`` / ``J`` are host copies (what the VTU holds)."""

    def __init__(self, mesh, dmesh: DeviceMesh, phi, J, stats, problem=None):
        self.mesh, self.dmesh, self.phi, self.J, self.stats, self.problem = mesh, dmesh, phi, J, stats, problem

    def close(self):
        self.dmesh.close()


_CTX = {}


def default_context(device=0) -> Context:
    if device not in _CTX:
        _CTX[device] = Context(device)
    return _CTX[device]


def solve_problem(ctx: Context, mesh, problem: sif.Problem, recover=DEFAULT_RECOVER, dmesh=None, **opts) -> SolvedCase:
    """Assemble + BCs + PCG + nodal current recovery for a parsed SIF problem on ``mesh``.
    ``dmesh`` lets a sweep reuse the device mesh/pattern (step04: same mesh, new conductivities)."""
    if not problem.dirichlet:
        raise ValueError("case has no Potential boundary condition (pure Neumann problem is singular)")
    dm = dmesh if dmesh is not None else ctx.mesh(mesh.nodes, mesh.tets, mesh.region, mesh.tris, mesh.bcid)
    dm.assemble(problem.sigma_by_body)
    dm.bc_reset(1)
    for bid, g in problem.neumann:
        dm.neumann(bid, g)
    for bid, v in problem.dirichlet:
        dm.dirichlet(bid, v)
    phi = dm.solve(**opts)[0]
    J = dm.recover_current(0, recover) if problem.calc_current else None
    return SolvedCase(mesh, dm, phi, J, dm.last_stats, problem)


def write_case_vtu(path, mesh, phi, J):
    pd = {"potential": np.ascontiguousarray(phi, dtype=np.float64)}
    if J is not None:
        pd["volume current"] = np.ascontiguousarray(J, dtype=np.float64)
    geom = np.concatenate([mesh.region.astype(np.int32), mesh.bcid.astype(np.int32)])
    vtu.write_vtu(path, mesh.nodes, mesh.tets, mesh.tris, point_data=pd, cell_data={"GeometryIds": geom})


def run_elmer_solver(run_dir, sif_name="case.sif", ctx=None, mesh=None, dmesh=None, recover=DEFAULT_RECOVER, keep=True, **opts):
    """In-process stand-in for ``subprocess.run(["ElmerSolver", "case.sif"], cwd=run_dir)``.
    Raises on any failure (the reference exits non-zero); returns the ``SolvedCase``."""
    run_dir = Path(run_dir)
    problem = sif.problem_from_sif((run_dir / sif_name).read_text())
    if mesh is None:
        mesh = elmer_io.read_elmer_mesh(run_dir / problem.mesh_db)
    ctx = ctx or default_context()
    case = solve_problem(ctx, mesh, problem, recover=recover, dmesh=dmesh, **opts)
    out_dir = run_dir / problem.results_dir
    out_dir.mkdir(parents=True, exist_ok=True)
    write_case_vtu(out_dir / f"{problem.output_name}_t0001.vtu", mesh, case.phi, case.J)
    if not keep:
        case.close()
    return case


# ---------------------------------------------------------------------------
# post-solve: metric rows (GPU reductions)
# ---------------------------------------------------------------------------
def _r(val, n):
    v = float(val)
    return round(v, n) if math.isfinite(v) else v


def _fp(pos, r, shape):
    return (float(pos[0]), float(pos[1]), float(r), shape == "square")


def skin_peaks(dm, z0_skin, t_skin, e1_pos, e2_pos, elec_r, shape, sys=0):
    """``run_layered_sweep.py:849-871``."""
    zmin = z0_skin + t_skin * 0.80
    a = dm.metric_nodes(0, zmin, sys=sys)
    if a["count"] == 0:
        return NAN, NAN
    peak_with = a["max"]
    b = dm.metric_nodes(0, zmin, mode=2, footprints=[_fp(e1_pos, elec_r, shape), _fp(e2_pos, elec_r, shape)], sys=sys)
    return peak_with, (b["max"] if b["count"] > 0 else peak_with)


def injected_current(dm, e1_pos, e2_pos, elec_r, z1, z2, shape, sys=0):
    """``run_layered_sweep.py:704-761``."""
    a = dm.metric_pad_current(z1 - max(z1 * 5e-3, 1e-5), _fp(e1_pos, elec_r, shape), 1.2, sys=sys)
    b = dm.metric_pad_current(z2 - max(z2 * 5e-3, 1e-5), _fp(e2_pos, elec_r, shape), 1.2, sys=sys)
    if a["count"] == 0 or b["count"] == 0:
        return (NAN,) * 5
    Ia, Ir = a["I_signed"], b["I_signed"]
    den = max(abs(Ia), abs(Ir))
    return abs(Ia), abs(Ir), (abs(Ia + Ir) / den if den > 0 else NAN), Ia, Ir


def compliance_voltage(dm, e1_pos, e2_pos, elec_r, z1, z2, shape, tol_from_active=False, sys=0):
    """``run_layered_sweep.py:899-920`` (step04 uses the active electrode's z tolerance for both
    masks and always disks: ``run_pressure_sweep.py:583-590`` -> ``tol_from_active``)."""
    t1 = max(z1 * 5e-3, 1e-5)
    t2 = t1 if tol_from_active else max(z2 * 5e-3, 1e-5)
    a = dm.metric_nodes(1, z1 - t1, mode=1, footprints=[_fp(e1_pos, elec_r, shape)], scale_r=1.5, sys=sys)
    if a["count"] == 0:
        return NAN
    b = dm.metric_nodes(1, z2 - t2, mode=1, footprints=[_fp(e2_pos, elec_r, shape)], scale_r=1.5, sys=sys)
    return a["sum"] / a["count"] - (b["sum"] / b["count"] if b["count"] > 0 else 0.0)


def eval_roi(dm, roi_cen, roi_radius_init, z0=0.0, z1=0.0, min_cells=4, include_tris=True, sys=0):
    """``run_layered_sweep.py:765-822``: returns (mean_J, mean_E, n, r_used, warning, fractions)."""
    mults = (1.0, 1.5, 2.0, 3.0)
    res = dm.metric_roi(roi_cen, roi_radius_init, mults, z0, z1, include_tris, sys=sys)
    pick, warning = None, None
    for mult, r in zip(mults, res):
        if r["n"] >= min_cells:
            pick = (mult, r)
            if mult > 1.0:
                warning = f"ROI radius expanded {mult:.1f}x to {roi_radius_init*mult*1000:.1f} mm ({r['n']} cells)"
            break
    if pick is None:
        pick = (3.0, res[-1])
        warning = f"ROI at 3x ({roi_radius_init*3*1000:.1f} mm) has only {res[-1]['n']} cells — noisy"
    mult, r = pick
    used = roi_radius_init * mult
    if r["n"] == 0:
        return NAN, NAN, 0, used, "No cells in ROI even at 3x expansion", (NAN, NAN, NAN)
    n = r["n"]
    frac = (r["n_above"] / n, r["n_mid"] / n, r["n_below"] / n)     # skin, fat, muscle
    return r["sum_J"] / n, r["sum_E"] / n, n, used, warning, frac


def extract_layered(case: SolvedCase, p, t_fat, elec_r, e1_pos, e2_pos, body_info, sigma_skin_used=None, jn_used=None,
                    elec_area_mesh=None, return_area_mesh=None, e1_id=None, e2_id=None, warn=print):
    """The 36-column step03 row (``run_layered_sweep.py:826-1030``; column order ``:991-1030``)."""
    dm = case.dmesh
    ls = p["layers"]
    st = p.get("stim", p.get("control", {}))
    z_skin_top = body_info["z_skin_top"]
    z1 = body_info.get("z_e1_elec_top", body_info["z_elec_top"])
    z2 = body_info.get("z_e2_elec_top", body_info["z_elec_top"])
    shape = body_info.get("elec_shape", "circle")
    t_sk = ls["t_skin"]
    peak_with, peak_no = skin_peaks(dm, z_skin_top - t_sk, t_sk, e1_pos, e2_pos, elec_r, shape)
    Ia, Ir, ferr, Ias, Irs = injected_current(dm, e1_pos, e2_pos, elec_r, z1, z2, shape)
    mode = st.get("control_mode", "voltage")
    comp, exceeded = NAN, False
    if mode == "current":
        I_target = st.get("injected_current_mA", 5.0) * 1e-3
        if math.isfinite(Ia) and I_target > 0 and abs(Ia - I_target) / I_target > 0.02:
            warn(f"    CURRENT ERROR: I_active={Ia*1e3:.3f} mA deviates {abs(Ia-I_target)/I_target:.1%} from "
                 f"target {I_target*1e3:.1f} mA (nodal-current integration, see README)")
        comp = compliance_voltage(dm, e1_pos, e2_pos, elec_r, z1, z2, shape)
        lim = st.get("compliance_voltage_V", 100.0)
        if math.isfinite(comp) and comp > lim:
            exceeded = True
            warn(f"    WARNING: compliance voltage {comp:.1f} V exceeds limit {lim:.0f} V")
    rc = p["roi"]
    z_nerve = z_skin_top - rc["z_target"]
    z_fat_bot, z_fat_top = z_skin_top - t_sk - t_fat, z_skin_top - t_sk
    mJ, mE, ncell, r_used, wmsg, (f_skin, f_fat, f_mus) = eval_roi(
        dm, [e1_pos[0], e1_pos[1], z_nerve], rc["roi_radius"], z0=z_fat_bot, z1=z_fat_top)
    if wmsg:
        warn(f"    WARNING: {wmsg}")
    area = math.pi * elec_r ** 2 if shape == "circle" else (2 * elec_r) ** 2
    eff = mE / peak_no if (math.isfinite(mE) and peak_no > 0) else NAN
    I_ref = Ia if math.isfinite(Ia) and Ia > 0 else NAN

    def norm(v):
        return v / I_ref if math.isfinite(v) and math.isfinite(I_ref) else NAN
    roi_layer = "skin" if z_nerve > z_fat_top else "fat" if z_nerve > z_fat_bot else "muscle"
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


def extract_pressure(case: SolvedCase, p, sigma_contact, label, e1_pos, e2_pos, body_info, jn_used, sys=0, warn=print):
    """The 24-column step04 row (``step04_pressure/run_pressure_sweep.py:528-660``)."""
    dm = case.dmesh
    ls, st = p["layers"], p.get("stim", p.get("control", {}))
    pl = p.get("placement", p.get("electrodes", {}))
    elec_r = float(pl["electrode_r_mm"]) * 1e-3
    shape = body_info["elec_shape"]
    z_skin_top = body_info["z_skin_top"]
    z1, z2 = body_info["z_e1_elec_top"], body_info["z_e2_elec_top"]
    t_skin = ls["t_skin"]
    peak_with, peak_no = skin_peaks(dm, z_skin_top - t_skin, t_skin, e1_pos, e2_pos, elec_r, "circle", sys=sys)
    Ia, Ir, ferr, Ias, Irs = injected_current(dm, e1_pos, e2_pos, elec_r, z1, z2, shape, sys=sys)
    comp = compliance_voltage(dm, e1_pos, e2_pos, elec_r, z1, z2, "circle", tol_from_active=True, sys=sys)
    lim = st.get("compliance_voltage_V", 200.0)
    exceeded = bool(math.isfinite(comp) and comp > lim)
    if exceeded:
        warn(f"    WARNING: compliance voltage {comp:.1f} V exceeds limit {lim:.0f} V")
    Z = comp / Ia if (math.isfinite(comp) and math.isfinite(Ia) and Ia > 0) else NAN
    rc = p["roi"]
    mJ, mE, ncell, r_used, wmsg, _ = eval_roi(dm, [e1_pos[0], e1_pos[1], z_skin_top - rc["z_target"]], rc["roi_radius"], sys=sys)
    if wmsg:
        warn(f"    WARNING: {wmsg}")
    pw_us = st.get("pulse_width_us", 200.0)
    charge = peak_with * pw_us * 1e-6 * 0.1 if math.isfinite(peak_with) else NAN
    limit = p.get("safety", {}).get("charge_density_limit_mC_cm2", 1.0)
    eff = mE / peak_no if (math.isfinite(mE) and peak_no > 0) else NAN
    return {
        "pressure_label": label, "sigma_contact_Spm": sigma_contact, "elec_r_mm": float(pl["electrode_r_mm"]),
        "t_fat_mm": ls["t_fat"] * 1000, "compliance_V": _r(comp, 3), "contact_impedance_ohm": _r(Z, 1),
        "exceeded_compliance": exceeded, "I_active_A": _r(Ia, 8), "I_return_A": _r(Ir, 8),
        "I_active_signed_A": _r(Ias, 8), "I_return_signed_A": _r(Irs, 8), "flux_err": _r(ferr, 6),
        "jn_used_A_m2": _r(jn_used, 6), "peak_J_skin_with_elec": _r(peak_with, 4),
        "peak_J_skin_no_elec": _r(peak_no, 4), "charge_density_mC_cm2": _r(charge, 6),
        "exceeds_charge_limit": bool(math.isfinite(charge) and charge > limit), "roi_mean_J": _r(mJ, 6),
        "roi_mean_E": _r(mE, 4), "efficiency": _r(eff, 6), "roi_n_cells": ncell,
        "roi_radius_used_mm": _r(r_used * 1000, 2), "pulse_width_us": pw_us,
        "frequency_Hz": st.get("frequency_Hz", 10.0),
    }


def extract_top_J(case: SolvedCase, Lz, sys=0):
    """``step02_electrodes/run_sweep.py:286-295,331-333``: (peak |J|, mean |J|, n) over nodes z > 0.99 Lz."""
    a = case.dmesh.metric_nodes(0, Lz * 0.99, sys=sys)
    if a["count"] == 0:
        return NAN, NAN, 0
    return a["max"], a["sum"] / a["count"], a["count"]


def step01_metrics(case: SolvedCase, sigma=0.2, v_top=1.0, v_bot=0.0):
    """``step01_box/test_step01_baseline.py:59-104``."""
    dm, pts = case.dmesh, case.mesh.nodes
    Lx, Ly, Lz = float(pts[:, 0].max()), float(pts[:, 1].max()), float(pts[:, 2].max())
    J_an = sigma * (v_top - v_bot) / Lz
    n, s1, _ = dm.metric_jstats(0.0)
    mean_J = s1 / n
    _, d1, d2 = dm.metric_jstats(mean_J)
    var = (d2 - d1 * d1 / n) / (n - 1) if n > 1 else 0.0
    std_J = math.sqrt(max(var, 0.0))
    m, sz, sf, szz, szf, sff = dm.metric_column_fit(Lx / 2, Ly / 2, Lx * 0.08)
    den = m * szz - sz * sz
    slope = (m * szf - sz * sf) / den
    icpt = (sf - slope * sz) / m
    ss_res = sff - 2 * slope * szf - 2 * icpt * sf + slope * slope * szz + 2 * slope * icpt * sz + m * icpt * icpt
    ss_tot = sff - sf * sf / m
    r2 = 1.0 - ss_res / ss_tot
    tol_z = Lz * 1e-3
    top = dm.metric_nodes(2, Lz - tol_z)
    bot = dm.metric_nodes(2, -1e300, zmax=tol_z)
    ft, fb = top["sum"] / top["count"], bot["sum"] / bot["count"]
    ph = dm.metric_nodes(1, -1e300)
    return dict(Lz=Lz, J_an=J_an, mean_J=mean_J, std_J=std_J, cv_J=std_J / mean_J, rel_J=abs(mean_J - J_an) / J_an,
                r2=r2, slope=slope, flux_top=ft, flux_bot=fb, flux_err=abs(ft - fb) / max(ft, fb),
                phi_min=ph["min"], phi_max=ph["max"])
