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
p.setdefault("mesh", {}).update({k: (int(v) if v.isdigit() else v) for k, v in opts.items()})
t_fat, r = 0.005, 0.010
with tempfile.TemporaryDirectory() as d:
    mesh, e1, e2, bi = s3.build_mesh(p, t_fat, r, Path(d) / "c", coarse=False)
    e1id, e2id, Aa, Ar = pipeline.detect_elec_bc_ids(mesh, e1, e2, e1[2], e2[2])
    jn = s3.write_sif(Path(d) / "c", e1id, e2id, p, r, bi, elec_area_mesh=Aa)
    prob = sif.problem_from_sif((Path(d) / "c" / "case.sif").read_text())
ref = fo.solve_case(mesh, prob.sigma_by_body, prob.dirichlet, prob.neumann, recover="lumped")
pts, J, phi = mesh.nodes, ref["J"], ref["phi"]
for name, bid, c in (("active", 101, e1), ("return", 102, e2)):
    tr = mesh.tris[mesh.bcid == bid]
    a = fo.tri_areas(pts, tr)
    jz = J[tr, 2].mean(axis=1)
    cen = pts[tr].mean(axis=1)
    d = np.hypot(cen[:, 0] - c[0], cen[:, 1] - c[1]) / r
    print(name, "I_nodal", (jz * a).sum(), "area", a.sum())
    for lo, hi in ((0, .5), (.5, .8), (.8, .9), (.9, 1.0)):
        m = (d >= lo) & (d < hi)
        print("   d/r %.1f-%.1f: area %.3e  mean Jz %.3f  contribution %.4e" % (lo, hi, a[m].sum(), (jz[m] * a[m]).sum() / a[m].sum(), (jz[m] * a[m]).sum()))
    # element-wise J_z of the pad tets under the top face (exact flux for comparison)
    el = fo.element_fields(pts, mesh.tets, mesh.region, prob.sigma_by_body, phi)
    Je = el[1] if isinstance(el, tuple) else el["J"]
    body = 4 if bid == 101 else 5
    m = mesh.region == body
    vol = meshgen.tet_volumes(pts, mesh.tets[m])
    print("   volume-mean element J_z in the pad x area:", (Je[m][:, 2] * vol).sum() / vol.sum() * a.sum())
