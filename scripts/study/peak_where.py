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
t_fat, r = 0.005, 0.010
with tempfile.TemporaryDirectory() as d:
    mesh, e1, e2, bi = s3.build_mesh(p, t_fat, r, Path(d) / "c", coarse=False)
    e1id, e2id, Aa, Ar = pipeline.detect_elec_bc_ids(mesh, e1, e2, e1[2], e2[2])
    jn = s3.write_sif(Path(d) / "c", e1id, e2id, p, r, bi, elec_area_mesh=Aa)
    prob = sif.problem_from_sif((Path(d) / "c" / "case.sif").read_text())
ref = fo.solve_case(mesh, prob.sigma_by_body, prob.dirichlet, prob.neumann, recover="lumped")
pts, J = mesh.nodes, ref["J"]
Jm = np.linalg.norm(J, axis=1)
top = np.abs(pts[:, 2] - 0.040) < 1e-9
print("jn", jn, "nodes on skin top", top.sum())
for name, c in (("active", e1), ("return", e2)):
    d = np.hypot(pts[:, 0] - c[0], pts[:, 1] - c[1])
    m = top & (d < 2 * r)
    idx = np.nonzero(m)[0]
    o = idx[np.argsort(-Jm[idx])][:8]
    print(name, "top-8 |J| nodes: ", [(round(d[i] / r, 3), round(Jm[i], 2), np.round(J[i], 1).tolist()) for i in o])
    bins = np.linspace(0, 1.6, 17)
    for a, b in zip(bins[:-1], bins[1:]):
        mm = m & (d / r >= a) & (d / r < b)
        if mm.any():
            print("  d/r %.1f-%.1f  n=%3d  mean %.2f  max %.2f" % (a, b, mm.sum(), Jm[mm].mean(), Jm[mm].max()))
