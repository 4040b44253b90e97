#!/usr/bin/env python3
"""step04: contact-pressure sweep (15 sigma_contact levels on ONE mesh) — drop-in for the reference's
``step04_pressure/run_pressure_sweep.py``: same CLI (``--smoke``), ``params.yaml`` schema, directory layout
(``results/_mesh_base``, ``results/p01..p15`` each with ``elmer_mesh/``, ``case.sif``,
``results/case_t0001.vtu``) and the 24-column ``summary.csv|json`` (``:635-660``).

The reference runs 15 separate ElmerSolver processes on a copied mesh (``:709-738``).  Here the mesh is
uploaded and its CSR pattern built once; by default all levels are assembled as a batch of matrices on
that pattern and solved by one batched-values PCG (K9), ``--sequential`` solves them one by one."""
import argparse
import csv
import json
import shutil
import sys
from pathlib import Path

import numpy as np
import yaml

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import _common  # noqa: F401,E402
from pelvistim_fem_b200 import elmer_io, pipeline, sif  # noqa: E402

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "step03_ankle_layers"))
import run_layered_sweep as step03  # noqa: E402

HERE = Path(__file__).resolve().parent
RESULTS_DIR = HERE / "results"
PARAMS_FILE = HERE / "params.yaml"
BATCH = 16   # systems per batched solve (library limit)


def load_params(path=PARAMS_FILE):
    with open(path) as f:
        return yaml.safe_load(f)


def build_mesh(p, run_dir, coarse=False):
    pl = p.get("placement", p.get("electrodes", {}))
    return step03.build_mesh(p, p["layers"]["t_fat"], float(pl["electrode_r_mm"]) * 1e-3, run_dir, coarse=coarse)


def run_pressure_sweep(p, sigma_contact_list, pressure_labels, coarse=False, sequential=False, ctx=None, results_dir=None):
    results_dir = Path(results_dir) if results_dir else RESULTS_DIR
    results_dir.mkdir(exist_ok=True)
    pl, st = p.get("placement", p.get("electrodes", {})), p.get("stim", p.get("control", {}))
    elec_r = float(pl["electrode_r_mm"]) * 1e-3
    print(f"\n{'='*60}")
    print("  PRESSURE SWEEP — sigma_contact vs compliance/charge/ROI")
    print(f"  Fixed: t_fat={p['layers']['t_fat']*1000:.0f}mm  r={pl['electrode_r_mm']:.0f}mm  "
          f"I={st['injected_current_mA']:.1f}mA  freq={st.get('frequency_Hz',10):.0f}Hz  pw={st.get('pulse_width_us',200):.0f}µs")
    print(f"  {len(sigma_contact_list)} pressure level(s): " +
          ", ".join(f"{lbl}({s:.4f})" for s, lbl in zip(sigma_contact_list, pressure_labels)))
    print(f"{'='*60}\n")
    mesh_dir = results_dir / "_mesh_base"
    print("  Building mesh (shared for all pressure levels)...")
    mesh, e1_pos, e2_pos, body_info = build_mesh(p, mesh_dir, coarse=coarse)
    print(f"    {mesh.nn} nodes")
    print("  Detecting electrode BCs ...")
    e1_id, e2_id, A_active, A_return = pipeline.detect_elec_bc_ids(mesh, e1_pos, e2_pos, e1_pos[2], e2_pos[2])
    print(f"    active={e1_id}  return={e2_id}  A_active={A_active*1e4:.4f}cm²  A_analytic={np.pi*elec_r**2*1e4:.4f}cm²")
    ctx = ctx or pipeline.default_context()
    dmesh = ctx.mesh(mesh.nodes, mesh.tets, mesh.region, mesh.tris, mesh.bcid)
    dmesh.pattern()
    # per level: case directory + SIF (what the reference hands to ElmerSolver)
    problems, jn_list = [], []
    for sigma_c, label in zip(sigma_contact_list, pressure_labels):
        run_dir = results_dir / label
        run_dir.mkdir(exist_ok=True)
        if (run_dir / "elmer_mesh").exists():
            shutil.rmtree(run_dir / "elmer_mesh")
        shutil.copytree(mesh_dir / "elmer_mesh", run_dir / "elmer_mesh")
        jn = step03.write_sif(run_dir, e1_id, e2_id, p, elec_r, body_info, elec_area_mesh=A_active,
                              sigma_contact_override=sigma_c, dialect="step04")
        (run_dir / "results").mkdir(exist_ok=True)
        problems.append(sif.problem_from_sif((run_dir / "case.sif").read_text()))
        jn_list.append(jn)
    all_results = []
    if sequential:
        for k, (sigma_c, label) in enumerate(zip(sigma_contact_list, pressure_labels)):
            print(f"\n[{label}]  sigma_contact={sigma_c:.4f} S/m")
            case = pipeline.run_elmer_solver(results_dir / label, ctx=ctx, mesh=mesh, dmesh=dmesh)
            all_results.append(_row(case, p, sigma_c, label, e1_pos, e2_pos, body_info, jn_list[k], 0))
    else:
        for b0 in range(0, len(problems), BATCH):
            chunk = list(range(b0, min(len(problems), b0 + BATCH)))
            print(f"\n  batched solve of levels {pressure_labels[chunk[0]]}..{pressure_labels[chunk[-1]]} "
                  f"({len(chunk)} matrices on one pattern)")
            dmesh.assemble([problems[k].sigma_by_body for k in chunk])
            dmesh.bc_reset(len(chunk))
            for j, k in enumerate(chunk):
                for bid, g in problems[k].neumann:
                    dmesh.neumann(bid, g, rhs=j)
                for bid, v in problems[k].dirichlet:
                    dmesh.dirichlet(bid, v, rhs=j)
            phi = dmesh.solve()
            for j, k in enumerate(chunk):
                label, sigma_c = pressure_labels[k], sigma_contact_list[k]
                J = dmesh.recover_current(j, "l2")
                case = pipeline.SolvedCase(mesh, dmesh, phi[j], J, dmesh.last_stats, problems[k])
                pipeline.write_case_vtu(results_dir / label / "results" / "case_t0001.vtu", mesh, phi[j], J)
                print(f"\n[{label}]  sigma_contact={sigma_c:.4f} S/m")
                all_results.append(_row(case, p, sigma_c, label, e1_pos, e2_pos, body_info, jn_list[k], j))
    dmesh.close()
    return all_results


def _row(case, p, sigma_c, label, e1_pos, e2_pos, body_info, jn, sys_idx):
    res = pipeline.extract_pressure(case, p, sigma_c, label, e1_pos, e2_pos, body_info, jn, sys=sys_idx)
    print(f"    compliance_V={res['compliance_V']:.1f} V  Z_contact={res['contact_impedance_ohm']:.0f} Ω  "
          f"charge={res['charge_density_mC_cm2']:.5f} mC/cm²  roi_E={res['roi_mean_E']:.2f} V/m")
    return res


def save_results(all_results, results_dir=None):
    step03.save_results(all_results, results_dir or RESULTS_DIR)


def main(argv=None):
    ap = argparse.ArgumentParser(description="Pressure-dependent contact sweep")
    ap.add_argument("--smoke", action="store_true", help="Single coarse case (middle pressure level)")
    ap.add_argument("--sequential", action="store_true", help="solve the levels one by one instead of as a batch")
    args = ap.parse_args(argv)
    p = load_params()
    ps = p["pressure_sweep"]
    sig_list, lbl_list = ps["sigma_contact_Spm"], ps["labels"]
    if args.smoke:
        mid = len(sig_list) // 2
        sig_list, lbl_list = [sig_list[mid]], [lbl_list[mid]]
        print("=== SMOKE TEST (1 coarse case) ===")
    results = run_pressure_sweep(p, sig_list, lbl_list, coarse=args.smoke, sequential=args.sequential)
    save_results(results)
    print(f"\n  {len(results)} pressure level(s) computed → results/summary.csv, results/summary.json")
    return results


if __name__ == "__main__":
    main()
