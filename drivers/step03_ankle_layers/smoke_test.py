#!/usr/bin/env python3
"""smoke_test.py — quick end-to-end validation for step03_ankle_layers (drop-in for the reference's
``step03_ankle_layers/smoke_test.py:81-188``: same checks, same order, same exit-code convention).

Runs a single coarse case (``run_layered_sweep.py --smoke``) and asserts:
  1. VTU output file exists
  2. Potential field: present, finite, range [0, 1] V in voltage mode / max > 0 in current mode
  3. Current density field ``volume current``: present, finite
  4. Electric field E = -grad(phi) computable and finite
  5. ``results/summary.json`` exists; current conservation at the electrode patches (flux_err < 5 %)
  6. total_current_A is positive and finite
  7. ROI mean |J| is positive and finite (cell-based, never NaN)
  8. compliance_V positive and finite (current mode)

The VTU is read with pyvista when it is installed (what the reference does, ``:88-123``) and with the package's own
reader otherwise; the solve behind ``--smoke`` is the GPU engine.

Usage (from step03_ankle_layers/):
    python3 smoke_test.py

Exit code 0 = all checks pass.  Non-zero = at least one failure (details printed).
"""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import yaml

HERE = Path(__file__).resolve().parent
PARAMS_FILE = HERE / "params.yaml"
RESULTS_DIR = HERE / "results"

FLUX_TOL = 0.05   # 5% tolerance for coarse mesh current conservation
ROI_MIN = 1e-6    # ROI mean |J| must exceed this (sanity floor)

PASS = "\033[32mPASS\033[0m"
FAIL = "\033[31mFAIL\033[0m"


def check(label, condition, detail=""):
    status = PASS if condition else FAIL
    line = f"  [{status}]  {label}"
    if detail:
        line += f"  ({detail})"
    print(line)
    return bool(condition)


def run_smoke_case():
    print("Running run_layered_sweep.py --smoke ...")
    r = subprocess.run([sys.executable, str(HERE / "run_layered_sweep.py"), "--smoke"], capture_output=False)
    if r.returncode != 0:
        print(f"\n{FAIL}  run_layered_sweep.py exited with code {r.returncode}")
        sys.exit(r.returncode)
    print()


def find_smoke_vtu(p):
    """Locate the VTU produced by the smoke case (configured fat thickness, middle electrode)."""
    t_fat = p["layers"]["t_fat"]
    pl = p.get("placement", p.get("electrodes", {}))
    r_list = pl.get("electrode_r_mm_list", pl.get("size_list", [5, 10, 15]))
    elec_r_mm = r_list[len(r_list) // 2]
    label = f"tfat{int(t_fat*1000):04d}um_r{int(elec_r_mm):04d}um"
    return RESULTS_DIR / label / "results" / "case_t0001.vtu", label, t_fat, elec_r_mm


def load_vtu(path):
    """points, point_data, tets - through pyvista when available, else the package's own reader."""
    try:
        import pyvista as pv
        mesh = pv.read(str(path))
        cells = mesh.cells_dict
        return np.array(mesh.points), {k: np.array(mesh.point_data[k]) for k in mesh.point_data.keys()}, np.array(cells.get(10, np.zeros((0, 4), int)))
    except ImportError:
        sys.path.insert(0, str(HERE.parent))
        import _common  # noqa: F401
        from pelvistim_fem_b200 import vtu
        v = vtu.read_vtu(path)
        tets, _ = vtu.split_cells(v)
        return v["points"], v["point_data"], tets


def cell_gradients(pts, tets, phi):
    """-grad(phi) of the linear interpolant in every tetrahedron (what ``compute_derivative`` evaluates at cell centres)."""
    p = pts[tets]
    d = p[:, 1:] - p[:, :1]                       # edge vectors from node 0
    dphi = phi[tets][:, 1:] - phi[tets][:, :1]
    return -np.linalg.solve(d, dphi[:, :, None])[:, :, 0]


def main():
    with open(PARAMS_FILE) as f:
        p = yaml.safe_load(f)
    run_smoke_case()
    vtu_path, label, t_fat, elec_r_mm = find_smoke_vtu(p)
    print(f"Checking case: {label}\n")
    passed = []
    # -- 1. VTU exists --------------------------------------------------------------------------------
    passed.append(check("VTU file exists", vtu_path.exists(), str(vtu_path)))
    if not vtu_path.exists():
        print("\nCannot continue — VTU not found.")
        sys.exit(1)
    pts, pd, tets = load_vtu(vtu_path)
    # -- 2. potential ---------------------------------------------------------------------------------
    phi_key = next((k for k in ("potential", "Potential") if k in pd), None)
    has_phi = phi_key is not None
    passed.append(check("Potential field present", has_phi, f"key='{phi_key}'"))
    mode_cfg = p.get("stim", p.get("control", {})).get("control_mode", "voltage")
    if has_phi:
        phi = np.array(pd[phi_key])
        passed.append(check("Potential is finite (no NaN/Inf)", np.all(np.isfinite(phi)), f"min={phi.min():.4f} max={phi.max():.4f} V"))
        if mode_cfg == "voltage":
            passed.append(check("Potential in [0, 1] V (voltage mode)", phi.min() >= -0.01 and phi.max() <= 1.01,
                                f"min={phi.min():.4f} max={phi.max():.4f}"))
        else:   # current mode: only the return electrode is grounded; the active one must be positive
            passed.append(check("Potential max > 0 V (current mode)", phi.max() > 0, f"min={phi.min():.4f} max={phi.max():.4f}"))
    # -- 3. current density ---------------------------------------------------------------------------
    has_J = "volume current" in pd
    passed.append(check("Field 'volume current' present", has_J))
    if has_J:
        Jmag = np.linalg.norm(np.array(pd["volume current"]), axis=1)
        passed.append(check("Current density is finite (no NaN/Inf)", np.all(np.isfinite(Jmag)), f"max|J|={Jmag.max():.3f} A/m²"))
    # -- 4. electric field ----------------------------------------------------------------------------
    if has_phi:
        try:
            E_cells = cell_gradients(pts, tets, phi)
            ok_E = E_cells.shape[0] > 0 and np.all(np.isfinite(E_cells))
        except Exception as exc:  # noqa: BLE001
            ok_E = False
            print(f"    E gradient error: {exc}")
        passed.append(check("E = -∇φ computable and finite", ok_E))
    # -- 5. summary.json + quantitative checks --------------------------------------------------------
    json_path = RESULTS_DIR / "summary.json"
    has_json = json_path.exists()
    passed.append(check("summary.json exists", has_json))
    if has_json and has_J:
        with open(json_path) as f:
            results = json.load(f)
        row = next((r for r in results if abs(r["t_fat_mm"] - t_fat * 1000) < 0.1 and abs(r["elec_r_mm"] - elec_r_mm) < 0.1), None)
        if row is not None:
            flux_err = row.get("flux_err", float("nan"))
            passed.append(check(f"Current conservation (flux_err < {FLUX_TOL:.0%})", np.isfinite(flux_err) and flux_err < FLUX_TOL,
                                f"flux_err = {flux_err:.3%}"))
            I_total = row.get("total_current_A", float("nan"))
            passed.append(check("total_current_A is positive and finite", np.isfinite(I_total) and I_total > 0,
                                f"total_current_A = {I_total:.4e} A"))
            roi_J = row.get("roi_mean_J", row.get("mean_J_roi", float("nan")))
            passed.append(check("ROI mean |J| is positive and finite", np.isfinite(roi_J) and roi_J > ROI_MIN,
                                f"roi_mean_J={roi_J:.5f} A/m²  roi_n_cells={row.get('roi_n_cells', 0)}  r_used={row.get('roi_radius_used_mm', '?')}mm"))
            if mode_cfg == "current":
                cV = row.get("compliance_V", float("nan"))
                lim = p.get("stim", {}).get("compliance_voltage_V", 100.0)
                passed.append(check("compliance_V is positive and finite", np.isfinite(cV) and cV > 0,
                                    f"compliance_V={cV:.2f} V  (limit={lim:.0f} V)"))
        else:
            print("  [SKIP]  Could not find matching row in summary.json")
    n_pass, n_total = sum(passed), len(passed)
    print(f"\n{'='*50}")
    print(f"Result: {n_pass}/{n_total} checks passed")
    if n_pass == n_total:
        print(f"[{PASS}]  All checks passed — pipeline is working.")
    else:
        print(f"[{FAIL}]  {n_total - n_pass} check(s) failed.")
        sys.exit(1)


if __name__ == "__main__":
    main()
