#!/usr/bin/env python3
"""CPU study (oracle only): step02 peak / mean |J| on the top face against the reference's PNG titles."""
import json, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from pelvistim_fem_b200 import meshgen, sizefield_mesher as sm
from oracle import fem_oracle as fo, metrics_oracle as mo
opts = dict(a.split("=") for a in sys.argv[1:])
mesher = opts.get("mesher", "graded"); recover = opts.get("recover", "lumped"); zf = float(opts.get("z_size_factor", 1.0))
gold = json.load(open(ROOT / "tests/golden/step02_png_titles.json"))
Lx, Ly, Lz, SEP = 0.15, 0.15, 0.05, 0.06
e1 = np.array([Lx / 2 - SEP / 2, Ly / 2]); e2 = np.array([Lx / 2 + SEP / 2, Ly / 2])
for shape in ("circle", "square"):
    for r in (0.005, 0.010, 0.015, 0.020):
        if mesher == "graded":
            m = sm.electrode_box_graded(Lx, Ly, Lz, e1, e2, r, shape, z_size_factor=zf)
        else:
            m = meshgen.electrode_box_mesh(Lx, Ly, Lz, e1, e2, r, shape, h_elec=r / 3.5, h_bulk=min(4 * r, 0.012), snap_rim=True)
        ref = fo.solve_case(m, {1: 0.2}, [(101, 1.0), (102, 0.0)], [], recover=recover)
        Jm = np.linalg.norm(ref["J"], axis=1)
        top = np.abs(m.nodes[:, 2] - Lz) < Lz * 1e-3
        pk, mn = gold[shape][str(int(round(r * 1000)))]
        print("%-7s r=%2d nn=%6d ntop=%5d  peak %7.2f (%+.3f)  mean %6.3f (%+.3f)" % (shape, r * 1000, m.nn, top.sum(), Jm[top].max(), Jm[top].max() / pk - 1, Jm[top].mean(), Jm[top].mean() / mn - 1), flush=True)
