#!/usr/bin/env python3
"""step02: electrode shape/size sweep on a homogeneous box (drop-in for the reference's
``step02_electrodes/run_sweep.py``: same constants ``:39-45``, labels ``:303``, case-directory layout
and printed peak/mean |J|; the plotting tail ``:346-480`` is out of scope).  Adds ``results/summary.csv|json``
(the reference keeps these numbers only in PNG titles).

Mesh: gmsh if importable, else the built-in structured mesher; ElmerGrid/ElmerSolver are replaced by
the in-process GPU engine."""
import argparse
import csv
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import _common  # noqa: F401,E402
from pelvistim_fem_b200 import elmer_io, gmsh_io, meshgen, pipeline, sif, sizefield_mesher  # noqa: E402

Lx, Ly, Lz = 0.15, 0.15, 0.05
SEP = 0.06
SIGMA = 0.2
SHAPES = ["circle", "square"]
RADII = [0.005, 0.010, 0.015, 0.020]
VOLTS = (1.0, 0.0)
RESULTS = Path("results")

e1_pos = np.array([Lx / 2 - SEP / 2, Ly / 2])
e2_pos = np.array([Lx / 2 + SEP / 2, Ly / 2])


def build_mesh(shape, r, run_dir, coarse=False, mesher="kuhn"):
    """``mesher``: "kuhn" (structured, default) or "graded" (size-field mesh of the reference's kind,
    ``sizefield_mesher.electrode_box_graded``: mean |J| over the top-face nodes within 2-5 % of the reference's PNG titles for
    r >= 10 mm against 13-20 %, but +34 % at r = 5 mm and square-corner peaks +22-27 % - oracle study, scripts/study/tables_cpu2.py)."""
    run_dir.mkdir(parents=True, exist_ok=True)
    s = 2.0 if coarse else 1.0
    if mesher == "graded":
        m = sizefield_mesher.electrode_box_graded(Lx, Ly, Lz, e1_pos, e2_pos, r, shape, lc_elec=s * r / 3.5, lc_bulk=s * min(4 * r, 0.012))
    else:
        m = meshgen.electrode_box_mesh(Lx, Ly, Lz, e1_pos, e2_pos, r, shape, h_elec=s * r / 3.5, h_bulk=s * min(4 * r, 0.012), snap_rim=True)
    gmsh_io.write_msh(run_dir / "mesh.msh", m, {(3, 1): "tissue", (2, 101): "active", (2, 102): "return", (2, 103): "other"})
    elmer_io.write_elmer_mesh(run_dir / "elmer_mesh", m)
    return m, (np.pi * r * r if shape == "circle" else (2 * r) ** 2)


def detect_elec_bc_ids(mesh):
    e1, e2, _, _ = pipeline.detect_elec_bc_ids(mesh, [*e1_pos, Lz], [*e2_pos, Lz], Lz, Lz)
    return e1, e2


def write_sif(run_dir, e1_id, e2_id):
    (run_dir / "case.sif").write_text(sif.serialize(sif.electrode_case(e1_id, e2_id, SIGMA, VOLTS[0], VOLTS[1])))


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--smoke", action="store_true", help="one coarse case")
    ap.add_argument("--mesher", choices=("kuhn", "graded"), default="kuhn")
    args = ap.parse_args(argv)
    RESULTS.mkdir(exist_ok=True)
    cases = [("circle", 0.010)] if args.smoke else [(s, r) for s in SHAPES for r in RADII]
    rows = []
    for shape, r in cases:
        label = f"{shape}_r{int(r*1000):02d}mm"
        run_dir = RESULTS / label
        print(f"\n[{label}]")
        print("  building mesh...")
        mesh, area = build_mesh(shape, r, run_dir, coarse=args.smoke, mesher=args.mesher)
        print("  detecting electrode boundary IDs...")
        e1_id, e2_id = detect_elec_bc_ids(mesh)
        print(f"    active BC={e1_id}, return BC={e2_id}")
        write_sif(run_dir, e1_id, e2_id)
        (run_dir / "results").mkdir(exist_ok=True)
        print("  running solver (GPU engine)...")
        case = pipeline.run_elmer_solver(run_dir, mesh=mesh)
        print("  extracting J on skin surface...")
        peak_J, mean_J, n_top = pipeline.extract_top_J(case, Lz)
        print(f"    peak|J|={peak_J:.2f}  mean|J|={mean_J:.2f} A/m²")
        rows.append(dict(shape=shape, r=r, area=area, label=label, n_nodes=mesh.nn, n_top_nodes=n_top,
                         peak_J=peak_J, mean_J=mean_J, iterations=case.stats["iterations"]))
        case.close()
    with open(RESULTS / "summary.csv", "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=list(rows[0].keys()))
        w.writeheader()
        w.writerows(rows)
    (RESULTS / "summary.json").write_text(json.dumps(rows, indent=2))
    print("\nAll simulations done.")
    return rows


if __name__ == "__main__":
    main()
