#!/usr/bin/env python3
"""CPU study (oracle only): step04 table columns against the reference's summary for mesher variants."""
import json, sys, tempfile
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "drivers")); sys.path.insert(0, str(ROOT / "drivers/step04_pressure")); sys.path.insert(0, str(ROOT / "drivers/step03_ankle_layers"))
import _common  # noqa
import run_pressure_sweep as s4
import run_layered_sweep as s3
from pelvistim_fem_b200 import pipeline, sif
from oracle import fem_oracle as fo, metrics_oracle as mo
opts = dict(a.split("=") for a in sys.argv[1:])
recover = opts.pop("recover", "lumped")
p = s4.load_params() if hasattr(s4, "load_params") else None
kw = {k: (int(v) if v.isdigit() else float(v) if v.replace(".", "").isdigit() else v) for k, v in opts.items()}
p.setdefault("mesh", {}).update(kw)
gold = json.load(open(ROOT / "tests/golden/step04_summary.json"))
pl = p.get("placement", p.get("electrodes", {}))
elec_r = float(pl["electrode_r_mm"]) * 1e-3
cols = ["compliance_V", "contact_impedance_ohm", "I_active_A", "I_return_A", "roi_mean_J", "roi_mean_E", "peak_J_skin_with_elec", "peak_J_skin_no_elec", "efficiency", "flux_err"]
print("opts", kw, "recover", recover)
print("%-6s" % "case" + "".join("%11s" % c[:10] for c in cols))
with tempfile.TemporaryDirectory() as d:
    mesh, e1, e2, bi = s4.build_mesh(p, Path(d) / "m")
    e1id, e2id, Aa, Ar = pipeline.detect_elec_bc_ids(mesh, e1, e2, e1[2], e2[2])
    worst = {c: 0.0 for c in cols}
    for g in gold:
        run = Path(d) / g["pressure_label"]; run.mkdir()
        jn = s3.write_sif(run, e1id, e2id, p, elec_r, bi, elec_area_mesh=Aa, sigma_contact_override=g["sigma_contact_Spm"], dialect="step04")
        prob = sif.problem_from_sif((run / "case.sif").read_text())
        ref = fo.solve_case(mesh, prob.sigma_by_body, prob.dirichlet, prob.neumann, recover=recover)
        row = mo.pressure_row(mesh.nodes, mesh.tets, mesh.tris, ref["phi"], ref["J"], p, g["sigma_contact_Spm"], g["pressure_label"], e1, e2, bi, jn)
        line = "%-6s" % g["pressure_label"]
        for c in cols:
            if c == "flux_err":
                line += "%6.3f/%.3f" % (row[c], g[c]); worst[c] = max(worst[c], row[c])
            else:
                rel = (row[c] - g[c]) / g[c]; line += "%+11.3f" % rel; worst[c] = max(worst[c], abs(rel))
        print(line, flush=True)
print("%-6s" % "worst" + "".join("%11.3f" % worst[c] for c in cols), "nn", mesh.nn)
