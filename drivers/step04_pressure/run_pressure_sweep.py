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
from pelvistim_fem_b200 import elmer_io, meshgen, pipeline, sif  # noqa: E402

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
    recover = p.get("solver", {}).get("current_recovery", pipeline.DEFAULT_RECOVER)
    all_results = []
    if sequential:
        for k, (sigma_c, label) in enumerate(zip(sigma_contact_list, pressure_labels)):
            print(f"\n[{label}]  sigma_contact={sigma_c:.4f} S/m")
            case = pipeline.run_elmer_solver(results_dir / label, ctx=ctx, mesh=mesh, dmesh=dmesh, recover=recover)
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
                J = dmesh.recover_current(j, recover)
                case = pipeline.SolvedCase(mesh, dmesh, phi[j], J, dmesh.last_stats, problems[k])
                pipeline.write_case_vtu(results_dir / label / "results" / "case_t0001.vtu", mesh, phi[j], J)
                print(f"\n[{label}]  sigma_contact={sigma_c:.4f} S/m")
                all_results.append(_row(case, p, sigma_c, label, e1_pos, e2_pos, body_info, jn_list[k], j))
    dmesh.close()
    return all_results


def run_compression_sweep(p, sigma_contact_list, pressure_labels, max_compression_mm, coarse=False, ctx=None, results_dir=None):
    """Pressure as conductivity AND geometry (BASELINE.json config 4; not in the reference, whose step04 changes
    sigma_contact only: ``run_pressure_sweep.py:12-13``): level k also indents the tissue under both pads by
    ``max_compression_mm * k/(n-1)``.  Topology is fixed, so the CSR pattern is built once; every level moves the
    nodes (``ptfem_mesh_set_coords``), re-assembles and solves.  Adds the column ``compression_mm``."""
    results_dir = Path(results_dir) if results_dir else RESULTS_DIR
    results_dir.mkdir(exist_ok=True)
    pl = p.get("placement", p.get("electrodes", {}))
    elec_r = float(pl["electrode_r_mm"]) * 1e-3
    mesh, e1_pos, e2_pos, body_info = build_mesh(p, results_dir / "_mesh_base", coarse=coarse)
    e1_id, e2_id, A_active, _ = pipeline.detect_elec_bc_ids(mesh, e1_pos, e2_pos, e1_pos[2], e2_pos[2])
    ctx = ctx or pipeline.default_context()
    dmesh = ctx.mesh(mesh.nodes, mesh.tets, mesh.region, mesh.tris, mesh.bcid)
    dmesh.pattern()
    Lz = p["geometry"]["Lz"]
    rows = []
    n = len(sigma_contact_list)
    for k, (sigma_c, label) in enumerate(zip(sigma_contact_list, pressure_labels)):
        depth = max_compression_mm * 1e-3 * (k / (n - 1) if n > 1 else 1.0)
        nodes_k = meshgen.compress_under_pads(mesh.nodes, [e1_pos[:2], e2_pos[:2]], elec_r, depth, Lz)
        mesh_k = type(mesh)(nodes_k, mesh.tets, mesh.region, mesh.tris, mesh.bcid, tri_parent=mesh.tri_parent)
        run_dir = results_dir / label
        run_dir.mkdir(exist_ok=True)
        elmer_io.write_elmer_mesh(run_dir / "elmer_mesh", mesh_k)
        z1 = float(nodes_k[np.unique(mesh.tris[mesh.bcid == e1_id]), 2].mean())
        z2 = float(nodes_k[np.unique(mesh.tris[mesh.bcid == e2_id]), 2].mean())
        bi = dict(body_info, z_e1_elec_top=z1, z_e2_elec_top=z2, z_elec_top=max(z1, z2))
        # pad areas change slightly with the indentation: Jn follows the deformed mesh, as write_sif does
        _, _, A_k, _ = pipeline.detect_elec_bc_ids(mesh_k, [*e1_pos[:2], z1], [*e2_pos[:2], z2], z1, z2)
        jn = step03.write_sif(run_dir, e1_id, e2_id, p, elec_r, bi, elec_area_mesh=A_k, sigma_contact_override=sigma_c,
                              dialect="step04")
        (run_dir / "results").mkdir(exist_ok=True)
        dmesh.set_coords(nodes_k)
        print(f"\n[{label}]  sigma_contact={sigma_c:.4f} S/m  compression={depth*1e3:.2f} mm")
        case = pipeline.run_elmer_solver(run_dir, ctx=ctx, mesh=mesh_k, dmesh=dmesh,
                                         recover=p.get("solver", {}).get("current_recovery", pipeline.DEFAULT_RECOVER))
        row = _row(case, p, sigma_c, label, [*e1_pos[:2], z1], [*e2_pos[:2], z2], bi, jn, 0)
        row["compression_mm"] = round(depth * 1e3, 4)
        rows.append(row)
    dmesh.close()
    return rows


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
    ap.add_argument("--compression-mm", type=float, default=0.0,
                    help="also indent the tissue under the pads, linearly up to this depth at the last level (geometry change "
                         "on fixed topology; extension, not in the reference)")
    args = ap.parse_args(argv)
    p = load_params()
    ps = p["pressure_sweep"]
    sig_list, lbl_list = ps["sigma_contact_Spm"], ps["labels"]
    if args.smoke:
        mid = len(sig_list) // 2
        sig_list, lbl_list = [sig_list[mid]], [lbl_list[mid]]
        print("=== SMOKE TEST (1 coarse case) ===")
    if args.compression_mm > 0:
        results = run_compression_sweep(p, sig_list, lbl_list, args.compression_mm, coarse=args.smoke)
    else:
        results = run_pressure_sweep(p, sig_list, lbl_list, coarse=args.smoke, sequential=args.sequential)
    save_results(results)
    print(f"\n  {len(results)} pressure level(s) computed → results/summary.csv, results/summary.json")
    return results


if __name__ == "__main__":
    main()
