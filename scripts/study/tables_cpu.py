#!/usr/bin/env python3
"""CPU study (oracle only): step03 table columns against the reference's summary for mesher variants.
usage: python scripts/study/tables_cpu.py [key=value ...]   e.g. n_skin=1 recover=lumped"""
import json, sys, time, tempfile
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "drivers")); sys.path.insert(0, str(ROOT / "drivers/step03_ankle_layers"))
import _common  # noqa
import run_layered_sweep as s3
from pelvistim_fem_b200 import pipeline, sif, meshgen
from oracle import fem_oracle as fo, metrics_oracle as mo

opts = dict(a.split("=") for a in sys.argv[1:])
recover = opts.pop("recover", "lumped")
mesh_kw = {k: (int(v) if v.lstrip("-").isdigit() else float(v) if v.replace(".", "").replace("e-", "").isdigit() else (v == "True") if v in ("True", "False") else v) for k, v in opts.items()}
p = s3.load_params()
gold = json.load(open(ROOT / "tests/golden/step03_summary.json"))
cols = ["elec_area_mesh_cm2", "roi_n_cells", "compliance_V", "total_current_A", "I_return_A", "roi_mean_J", "roi_mean_E", "peak_J_skin_with_elec", "peak_J_skin_no_elec", "efficiency", "flux_err"]
p.setdefault("mesh", {}).update(mesh_kw)          # mesher=graded|kuhn n_skin=.. n_fat=..
print("opts", mesh_kw, "recover", recover)
print("%-14s" % "case" + "".join("%11s" % c[:10] for c in cols))
worst = {c: 0.0 for c in cols}
for g in gold:
    t_fat, r = g["t_fat_mm"] * 1e-3, g["elec_r_mm"] * 1e-3
    with tempfile.TemporaryDirectory() as d:
        mesh, e1, e2, bi = s3.build_mesh(p, t_fat, r, Path(d) / "c", coarse=False)
        e1id, e2id, Aa, Ar = pipeline.detect_elec_bc_ids(mesh, e1, e2, e1[2], e2[2])
        jn = s3.write_sif(Path(d) / "c", e1id, e2id, p, r, bi, elec_area_mesh=Aa)
        prob = sif.problem_from_sif((Path(d) / "c" / "case.sif").read_text())
    ref = fo.solve_case(mesh, prob.sigma_by_body, prob.dirichlet, prob.neumann, recover=recover)
    row = mo.layered_row(mesh.nodes, mesh.tets, mesh.tris, ref["phi"], ref["J"], p, t_fat, r, e1, e2, bi, jn_used=jn,
                         elec_area_mesh=Aa, return_area_mesh=Ar, e1_id=e1id, e2_id=e2id)
    line = "%-14s" % ("t%g_r%g n=%d" % (g["t_fat_mm"], g["elec_r_mm"], mesh.nn))
    for c in cols:
        if c == "flux_err":
            line += "%6.3f/%.3f" % (row[c], g[c]); worst[c] = max(worst[c], row[c])
        else:
            rel = (row[c] - g[c]) / g[c]; line += "%+11.3f" % rel; worst[c] = max(worst[c], abs(rel))
    print(line, flush=True)
print("%-14s" % "worst" + "".join("%11.3f" % worst[c] for c in cols))
