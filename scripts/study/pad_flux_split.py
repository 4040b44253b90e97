import json, sys, tempfile
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "drivers")); sys.path.insert(0, str(ROOT / "drivers/step03_ankle_layers"))
import _common  # noqa
import run_layered_sweep as s3
from pelvistim_fem_b200 import pipeline, sif, meshgen
from oracle import fem_oracle as fo, metrics_oracle as mo
p = s3.load_params()
opts = dict(a.split("=") for a in sys.argv[1:])
coarse = opts.pop("coarse", "0") == "1"
recover = opts.pop("recover", "lumped")
p.setdefault("mesh", {}).update({k: (int(v) if v.isdigit() else float(v) if v.replace(".","").isdigit() else v) for k, v in opts.items()})
t_fat, r = 0.005, 0.010
with tempfile.TemporaryDirectory() as d:
    mesh, e1, e2, bi = s3.build_mesh(p, t_fat, r, Path(d) / "c", coarse=coarse)
    e1id, e2id, Aa, Ar = pipeline.detect_elec_bc_ids(mesh, e1, e2, e1[2], e2[2])
    jn = s3.write_sif(Path(d) / "c", e1id, e2id, p, r, bi, elec_area_mesh=Aa)
    prob = sif.problem_from_sif((Path(d) / "c" / "case.sif").read_text())
ref = fo.solve_case(mesh, prob.sigma_by_body, prob.dirichlet, prob.neumann, recover=recover)
pts, J = mesh.nodes, ref["J"]
tr = mesh.tris
a = fo.tri_areas(pts, tr); jz = J[tr, 2].mean(axis=1); cen = pts[tr].mean(axis=1)
z_top = e1[2]; tol_z = max(z_top * 5e-3, 1e-5)
for name, c in (("active", e1), ("return", e2)):
    d = np.hypot(cen[:, 0] - c[0], cen[:, 1] - c[1])
    m = (cen[:, 2] > z_top - tol_z) & (d < 1.2 * r)
    flat = m & (np.abs(pts[tr][:, :, 2] - z_top).max(axis=1) < 1e-9)
    wall = m & ~flat
    print(name, "nn", mesh.nn, "top-face %.4f mA (%d tris)  wall %.4f mA (%d tris, area %.2f mm2)  total %.4f" % ((jz[flat] * a[flat]).sum() * 1e3, flat.sum(), (jz[wall] * a[wall]).sum() * 1e3, wall.sum(), a[wall].sum() * 1e6, (jz[m] * a[m]).sum() * 1e3))
row = mo.layered_row(mesh.nodes, mesh.tets, mesh.tris, ref["phi"], ref["J"], p, t_fat, r, e1, e2, bi, jn_used=jn, elec_area_mesh=Aa, return_area_mesh=Ar, e1_id=e1id, e2_id=e2id)
print("flux_err", row["flux_err"], "Ia", row["total_current_A"], "Ir", row["I_return_A"], "peak_with", row["peak_J_skin_with_elec"])
