"""step03 (9 rows) + step04 (rows p01, p08, p15) through the ORACLE on CPU for given `mesh:` overrides, every column of the GPU driver
tests against the reference's committed tables with those tests' tolerances (BAD marks a miss).  Generator of
profiles/r03_tables_cpu_oracle.txt.
    python scripts/study/tables_cpu_roi.py '{"z_size_factor": 1.25, "z_volume_law": False}' ['(5, 8)']   # overrides, t_fat rows in mm"""
import sys, time, tempfile, json, copy
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
for q in ("", "drivers/step04_pressure", "drivers/step03_ankle_layers", "drivers"):
    sys.path.insert(0, str(ROOT / q))
import numpy as np
import pelvistim_fem_b200
from pelvistim_fem_b200 import pipeline, sif
from oracle import fem_oracle as fo, metrics_oracle as mo
import run_layered_sweep as s3, run_pressure_sweep as s4
from pathlib import Path
over = eval(sys.argv[1]) if len(sys.argv) > 1 else {}
G3 = json.load(open(ROOT / "tests/golden/step03_summary.json"))
G4 = json.load(open(ROOT / "tests/golden/step04_summary.json"))
T3 = (("compliance_V", 0.015), ("total_current_A", 0.02), ("I_return_A", 0.08), ("roi_mean_J", 0.07), ("roi_mean_E", 9), ("peak_J_skin_with_elec", 0.05), ("peak_J_skin_no_elec", 0.25), ("efficiency", 9))
T4 = (("compliance_V", 0.015), ("contact_impedance_ohm", 0.015), ("I_active_A", 0.003), ("I_return_A", 0.025), ("roi_mean_J", 0.04), ("roi_mean_E", 0.14), ("peak_J_skin_with_elec", 0.15), ("peak_J_skin_no_elec", 0.17), ("efficiency", 9))
def solve(mesh, run_dir):
    prob = sif.problem_from_sif((run_dir / "case.sif").read_text())
    return fo.solve_case(mesh, prob.sigma_by_body, prob.dirichlet, prob.neumann, recover=pipeline.DEFAULT_RECOVER)
p = s3.load_params(); p.setdefault("mesh", {}).update(over)
k = 0
only = eval(sys.argv[2]) if len(sys.argv) > 2 else None
for t_fat in p["layers"]["t_fat_sweep"]:
    for r_mm in p["placement"]["electrode_r_mm_list"]:
        elec_r = r_mm * 1e-3
        if only is not None and round(t_fat * 1000) not in only:
            k += 1
            continue
        with tempfile.TemporaryDirectory() as d:
            rd = Path(d) / "c"
            mesh, e1, e2, bi = s3.build_mesh(p, t_fat, elec_r, rd)
            e1id, e2id, Aa, Ar = pipeline.detect_elec_bc_ids(mesh, e1, e2, e1[2], e2[2])
            jn = s3.write_sif(rd, e1id, e2id, p, elec_r, bi, elec_area_mesh=Aa)
            ref = solve(mesh, rd)
        row = mo.layered_row(mesh.nodes, mesh.tets, mesh.tris, ref["phi"], ref["J"], p, t_fat, elec_r, e1, e2, bi, jn_used=jn, elec_area_mesh=Aa, return_area_mesh=Ar, e1_id=e1id, e2_id=e2id)
        g = G3[k]; k += 1
        out = {c: round((row[c] - g[c]) / abs(g[c]), 3) for c, _ in T3}
        bad = [c for c, tol in T3 if abs(row[c] - g[c]) > tol * abs(g[c])]
        print("s3", g["t_fat_mm"], g["elec_r_mm"], "nn", mesh.nn, "roi_n", row["roi_n_cells"], g["roi_n_cells"], "flux", row["flux_err"], out, "BAD" if bad or row["flux_err"] >= 0.045 else "", bad, flush=True)
p = s4.load_params(); p.setdefault("mesh", {}).update(over)
ps = p["pressure_sweep"]
with tempfile.TemporaryDirectory() as d:
    mesh, e1, e2, bi = s4.build_mesh(p, Path(d) / "m")
    e1id, e2id, Aa, Ar = pipeline.detect_elec_bc_ids(mesh, e1, e2, e1[2], e2[2])
    import inspect
    for idx in ((7,) if only is not None else (0, 7, 14)):
        sc, lbl = ps["sigma_contact_Spm"][idx], ps["labels"][idx]
        rd = Path(d) / lbl; rd.mkdir()
        jn = s4.step03.write_sif(rd, e1id, e2id, p, float(p["placement"]["electrode_r_mm"]) * 1e-3, bi, elec_area_mesh=Aa,
                                 sigma_contact_override=sc, dialect="step04")
        ref = solve(mesh, rd)
        row = mo.pressure_row(mesh.nodes, mesh.tets, mesh.tris, ref["phi"], ref["J"], p, sc, lbl, e1, e2, bi, jn)
        g = G4[idx]
        out = {c: round((row[c] - g[c]) / abs(g[c]), 3) for c, _ in T4}
        bad = [c for c, tol in T4 if abs(row[c] - g[c]) > tol * abs(g[c])]
        print("s4", lbl, "roi_n", row["roi_n_cells"], g["roi_n_cells"], "flux", row["flux_err"], out, "BAD" if bad or row["flux_err"] >= 0.03 else "", bad, flush=True)
