#!/usr/bin/env python3
"""step03: layered ankle slab sweep (fat thickness x electrode radius) — drop-in for the reference's
``step03_ankle_layers/run_layered_sweep.py``: same CLI (``--smoke``), ``params.yaml`` schema, case labels
(``:1063-1064``), per-case files (``elmer_mesh/``, ``case.sif``, ``bc_debug_report.txt``,
``results/case_t0001.vtu``) and ``results/summary.csv|json`` columns (``:991-1030``).

What changed: ``gmsh`` meshing falls back to the built-in structured mesher, and the two subprocess
boundaries (``ElmerGrid`` ``:1077``, ``ElmerSolver`` ``:1099``) plus the pyvista metric extraction are
served by the GPU engine in-process.  Extra flags: ``--gpus N`` shards the independent sweep points
over N GPUs (one worker process per GPU; rows are gathered in sweep order)."""
import argparse
import csv
import json
import math
import sys
from pathlib import Path

import numpy as np
import yaml

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import _common  # noqa: F401,E402
from pelvistim_fem_b200 import elmer_io, gmsh_io, meshgen, pipeline, sif, sizefield_mesher, sweep  # noqa: E402

HERE = Path(__file__).resolve().parent
RESULTS_DIR = HERE / "results"
PARAMS_FILE = HERE / "params.yaml"
_PLANS = {}          # 2-D triangulations of the graded mesher, by pad geometry


def load_params(path=PARAMS_FILE):
    with open(path) as f:
        return yaml.safe_load(f)


def _pl(p):
    return p.get("placement", p.get("electrodes", {}))


def _stim(p):
    return p.get("stim", p.get("control", {}))


def build_mesh(p, t_fat, elec_r, run_dir, coarse=False):
    """Layered slab + contact pads (``run_layered_sweep.py:122-362``, rect cross-section).
    Returns ``(mesh, e1_pos3d, e2_pos3d, body_info)``; writes ``run_dir/elmer_mesh``."""
    run_dir = Path(run_dir)
    run_dir.mkdir(parents=True, exist_ok=True)
    g, ls, pl = p["geometry"], p["layers"], _pl(p)
    Lx, Ly, Lz = g["Lx"], g["Ly"], g["Lz"]
    if g.get("cross_section", "rect") != "rect":
        raise NotImplementedError("only cross_section: rect (the configured default) is built in")
    shape = pl.get("electrode_shape", pl.get("shape", "circle"))
    active_xy = pl.get("active_xy", [pl.get("medial_offset", 0.025), Ly / 2])
    return_xy = pl.get("return_xy", [Lx - pl.get("lateral_offset", 0.025), Ly / 2])
    ct = p.get("contact", {})
    contact = bool(ct.get("enabled", False))
    t_contact = ct.get("t_contact_mm", 0.5) * 1e-3 if contact else 0.0
    mm = p.get("mesh", {})
    scale = 2.0 if coarse else 1.0
    lc_elec = mm.get("lc_electrode_mm", elec_r * 300) * 1e-3 * scale
    lc_bulk = mm.get("lc_global_mm", 3.0) * 1e-3 * scale
    t_muscle = Lz - ls["t_skin"] - t_fat
    bn = p.get("bone", {})
    bone = None
    if bn.get("enabled", False):   # optional bone block inside the muscle (params.yaml `bone:`; not in the reference's model)
        bone = dict(x=[v * 1e-3 for v in bn["x_mm"]], y=[v * 1e-3 for v in bn["y_mm"]], z=[v * 1e-3 for v in bn["z_mm"]])
    # (the rim of the footprint is snapped onto the circle on the production meshes only: on the doubled spacing of the smoke
    #  case the snapped rim elements are distorted enough to cost accuracy - pad-current mismatch 5.1 % against 4.3 % unsnapped,
    #  3.0 % against 4.1 % at full resolution; oracle study)
    mesher = mm.get("mesher", "graded")
    if mesher == "graded":
        # size-field mesh of the reference's kind (Distance/Threshold law, polygonal pad rim, run_layered_sweep.py:311-323);
        # the 2-D triangulation depends on the pad size only and is kept between sweep points
        key = (Lx, Ly, tuple(active_xy), tuple(return_xy), elec_r, shape, lc_elec, lc_bulk)
        mesh = sizefield_mesher.layered_slab_graded(
            Lx, Ly, Lz, ls["t_skin"], t_fat, t_contact or 0.0005, active_xy, return_xy, elec_r, shape, lc_elec=lc_elec,
            lc_bulk=lc_bulk, n_skin=mm.get("n_skin"), n_fat=mm.get("n_fat"), n_contact=mm.get("n_contact", 1),
            contact_enabled=contact, bone=bone, z_size_factor=float(mm.get("z_size_factor", 1.5)),
            z_volume_law=bool(mm.get("z_volume_law", True)), z_size_factor_fat=float(mm.get("z_size_factor_fat", 1.25)),
            plan=_PLANS.get(key))
        _PLANS[key] = mesh.meta["plan"]
    elif mesher == "kuhn":
        n_m = max(3, int(round(t_muscle / lc_bulk)))
        n_f = max(2, int(round(t_fat / lc_elec)))
        mesh = meshgen.layered_slab_mesh(Lx, Ly, Lz, ls["t_skin"], t_fat, t_contact or 0.0005, active_xy, return_xy, elec_r,
                                         shape, n_muscle=n_m, n_fat=n_f, n_skin=2, n_contact=1, h_bulk=lc_bulk,
                                         h_elec=lc_elec, contact_enabled=contact, snap_rim=not coarse, bone=bone)
    else:
        raise ValueError(f"mesh.mesher: {mesher!r} (graded | kuhn)")
    # same per-case files as the reference: mesh.msh (gmsh.write, :342-343) then the ElmerGrid 14 2 conversion (:1077)
    names = {(3, 1): "muscle", (3, 2): "fat", (3, 3): "skin", (3, 4): "contact_active", (3, 5): "contact_return", (3, 6): "bone",
             (2, 101): "active", (2, 102): "return", (2, 103): "other"}
    gmsh_io.write_msh(run_dir / "mesh.msh", mesh, names)
    elmer_io.write_elmer_mesh(run_dir / "elmer_mesh", mesh)
    z_top = Lz + t_contact
    body_info = dict(contact_enabled=contact, z_skin_top=Lz, z_elec_top=z_top, z_e1_skin=Lz, z_e2_skin=Lz,
                     z_e1_elec_top=z_top, z_e2_elec_top=z_top, c1_body_id=4 if contact else None,
                     c2_body_id=5 if contact else None, elec_shape=shape, bone=mesh.meta.get("bone"))
    e1 = np.array([float(active_xy[0]), float(active_xy[1]), z_top])
    e2 = np.array([float(return_xy[0]), float(return_xy[1]), z_top])
    return mesh, e1, e2, body_info


def write_sif(run_dir, e1_id, e2_id, p, elec_r, body_info, sigma_skin_override=None, elec_area_mesh=None,
              sigma_contact_override=None, dialect="step03"):
    c, sv, st, ct = p["conductivities"], p["solver"], _stim(p), p.get("contact", {})
    sigma_skin = sigma_skin_override if sigma_skin_override is not None else c["sigma_skin"]
    sigma_c = sigma_contact_override if sigma_contact_override is not None else ct.get("sigma_contact_Spm", 0.005)
    secs, jn_used, warning = sif.layered_case(
        e1_id, e2_id, c["sigma_muscle"], c["sigma_fat"], sigma_skin, sigma_c,
        contact=body_info.get("contact_enabled", False), c1_body=body_info.get("c1_body_id") or 4,
        c2_body=body_info.get("c2_body_id") or 5, mode=st.get("control_mode", "voltage"),
        injected_current_mA=st.get("injected_current_mA", 5.0), elec_r=elec_r, shape=body_info.get("elec_shape", "circle"),
        elec_area_mesh=elec_area_mesh, tol=sv.get("tolerance", 1e-8), lin_solver=sv.get("linear_solver", "UMFPACK"),
        dialect=dialect, sigma_bone=(c.get("sigma_bone", 0.02) if body_info.get("bone") else None))
    if warning:
        print(f"    WARNING: {warning}")
    (Path(run_dir) / "case.sif").write_text(sif.serialize(secs))
    return jn_used


def nerve_polyline(p, e1_pos, body_info):
    """Sample points along the (straight) tibial-nerve fibre: by default parallel to x at the ROI depth
    (``roi.z_target`` below the skin, ``params.yaml:72-74``) under the electrode row, 5 mm clear of the slab ends.
    Optional ``nerve: {points: [[x,y,z],...], internode_mm: h}`` in params.yaml gives an explicit polyline and
    sample spacing.  The spacing defaults to 2 mm (node-of-Ranvier scale, and not below the element size: the
    second difference of a P1 potential sampled finer than the mesh is a train of kinks)."""
    nv = p.get("nerve", {})
    h = float(nv.get("internode_mm", 2.0)) * 1e-3
    if "points" in nv:
        ctrl = np.asarray(nv["points"], dtype=float)
    else:
        g = p["geometry"]
        z = body_info["z_skin_top"] - p["roi"]["z_target"]
        ctrl = np.array([[0.005, e1_pos[1], z], [g["Lx"] - 0.005, e1_pos[1], z]])
    seg = np.linalg.norm(np.diff(ctrl, axis=0), axis=1)
    s_ctrl = np.concatenate([[0.0], np.cumsum(seg)])
    n = max(3, int(round(s_ctrl[-1] / h)) + 1)
    s = np.linspace(0.0, s_ctrl[-1], n)
    pts = np.stack([np.interp(s, s_ctrl, ctrl[:, k]) for k in range(3)], axis=1)
    return s, pts


def save_activating_function(case, p, e1_pos, body_info, run_dir):
    """Potential and activating function (second difference of phi along the fibre, V/m^2) sampled on the GPU
    (K13; BASELINE.json north_star - the reference only has the ROI-mean |E| proxy, ``run_layered_sweep.py:930-936``)."""
    s, pts = nerve_polyline(p, e1_pos, body_info)
    phi, af = case.dmesh.sample_polyline(pts)
    out = Path(run_dir) / "results" / "activating_function.csv"
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["s_m", "x_m", "y_m", "z_m", "phi_V", "activating_function_V_per_m2"])
        for k in range(len(s)):
            w.writerow([f"{s[k]:.6e}", f"{pts[k,0]:.6e}", f"{pts[k,1]:.6e}", f"{pts[k,2]:.6e}", f"{phi[k]:.9e}", f"{af[k]:.9e}"])
    ok = np.isfinite(af)
    return float(np.max(af[ok])) if ok.any() else float("nan"), out


def case_label(t_fat, elec_r):
    return f"tfat{int(t_fat*1000):04d}um_r{int(elec_r*1000):04d}um"


def displace_fat_thickness(mesh, p, t_fat_ref, t_fat):
    """The mesh built for fat thickness ``t_fat_ref`` with its nodes moved so that the fat layer is ``t_fat`` thick - same
    topology, piecewise-linear map of z: the muscle block is stretched, the fat layer squeezed, skin and pads stay put.
    (The reference re-meshes every point, ``run_layered_sweep.py:1061-1062``; on fixed topology the CSR pattern, the
    element -> non-zero map and the device mesh are built once per electrode size.)"""
    Lz, t_skin = p["geometry"]["Lz"], p["layers"]["t_skin"]
    z_fs = Lz - t_skin
    z_ref, z_new = z_fs - t_fat_ref, z_fs - t_fat
    if z_new <= 1e-4:
        raise ValueError(f"t_muscle = {z_new*1000:.2f} mm <= 0.1 mm - reduce t_fat + t_skin or increase Lz")
    nodes = mesh.nodes.copy()
    z = nodes[:, 2]
    nodes[:, 2] = np.where(z <= z_ref, z * (z_new / z_ref), np.where(z <= z_fs, z_new + (z - z_ref) * (t_fat / t_fat_ref), z))
    return type(mesh)(nodes, mesh.tets, mesh.region, mesh.tris, mesh.bcid)


def run_case(p, t_fat, elec_r, coarse=False, sigma_skin_override=None, ctx=None, results_dir=None, quiet=False, prebuilt=None,
             dmesh=None):
    """One sweep point, end to end (the loop body of ``run_layered_sweep.py:1061-1124``).  ``prebuilt`` =
    ``(mesh, e1_pos, e2_pos, body_info)`` and ``dmesh`` (a device mesh of the same topology whose coordinates have been set
    to ``mesh.nodes``): the fixed-topology sweep, which re-uses pattern and device mesh instead of re-meshing."""
    say = (lambda *a, **k: None) if quiet else print
    results_dir = Path(results_dir) if results_dir else RESULTS_DIR
    sigma_skin = sigma_skin_override if sigma_skin_override is not None else p["conductivities"]["sigma_skin"]
    label = case_label(t_fat, elec_r)
    run_dir = results_dir / label
    say(f"\n[{label}]  t_fat={t_fat*1000:.1f}mm  r={elec_r*1000:.1f}mm  sigma_skin={sigma_skin}")
    if prebuilt is None:
        say("  meshing ...")
        mesh, e1_pos, e2_pos, body_info = build_mesh(p, t_fat, elec_r, run_dir, coarse=coarse)
    else:
        say("  mesh: fixed topology, nodes displaced ...")
        mesh, e1_pos, e2_pos, body_info = prebuilt
        run_dir.mkdir(parents=True, exist_ok=True)
        elmer_io.write_elmer_mesh(run_dir / "elmer_mesh", mesh)
    say(f"    {mesh.nn} nodes")
    say("  detecting electrode BCs + computing mesh areas ...")
    e1_id, e2_id, A_act, A_ret = pipeline.detect_elec_bc_ids(mesh, e1_pos, e2_pos, e1_pos[2], e2_pos[2])
    shape = body_info["elec_shape"]
    area_an = math.pi * elec_r ** 2 if shape == "circle" else (2 * elec_r) ** 2
    say(f"    active={e1_id}  return={e2_id}  A_active={A_act*1e4:.4f}cm²  A_analytic={area_an*1e4:.4f}cm²")
    jn_used = write_sif(run_dir, e1_id, e2_id, p, elec_r, body_info, sigma_skin_override=sigma_skin_override,
                        elec_area_mesh=A_act)
    pipeline.save_bc_debug_report(run_dir, label, e1_id, e2_id, A_act, A_ret, jn_used, p, body_info)
    (run_dir / "results").mkdir(exist_ok=True)
    say("  solver (GPU engine) ...")
    case = pipeline.run_elmer_solver(run_dir, ctx=ctx, mesh=mesh, dmesh=dmesh,
                                     recover=p.get("solver", {}).get("current_recovery", pipeline.DEFAULT_RECOVER))
    say("  extracting metrics ...")
    res = pipeline.extract_layered(case, p, t_fat, elec_r, e1_pos, e2_pos, body_info, sigma_skin_used=sigma_skin,
                                   jn_used=jn_used, elec_area_mesh=A_act, return_area_mesh=A_ret, e1_id=e1_id,
                                   e2_id=e2_id, warn=say)
    af_peak, af_path = save_activating_function(case, p, e1_pos, body_info, run_dir)
    say(f"    activating function along the nerve fibre: peak {af_peak:.4e} V/m² → {af_path.name}")
    if dmesh is None:
        case.close()
    say(f"    peak_J_no_elec={res['peak_J_skin_no_elec']:.4f}  roi_mean_E={res['roi_mean_E']:.4f}  "
        f"efficiency={res['efficiency']:.4e}  flux_err={res['flux_err']:.3e}")
    if res.get("control_mode") == "current":
        say(f"    compliance_V={res['compliance_V']:.2f} V  I_active={res['total_current_A']:.4e} A  "
            f"I_return={res['I_return_A']:.4e} A")
    return res


def _point(args):
    p, t_fat, elec_r, coarse, override, results_dir = args
    return run_case(p, t_fat, elec_r, coarse, override, ctx=sweep.worker_context(), results_dir=results_dir,
                    quiet=sweep.worker_rank() is not None)


def run_sweep(p, t_fat_list, elec_r_list, coarse=False, sigma_skin_override=None, gpus=1, pipelines=1):
    RESULTS_DIR.mkdir(exist_ok=True)
    st = _stim(p)
    mode = st.get("control_mode", "voltage")
    print(f"\n{'='*60}")
    if mode == "current":
        print("  CONTROL MODE : current")
        print(f"  Injected I   : {st.get('injected_current_mA', 5.0):.1f} mA  (per-case Neumann BC at active electrode)")
        print(f"  Compliance   : warn if V_active > {st.get('compliance_voltage_V', 100.0):.0f} V")
    else:
        print("  CONTROL MODE : voltage")
        print("  V_active = 1.0 V  |  V_return = 0 V  (Dirichlet BCs)")
    print(f"{'='*60}\n")
    points = [(p, t_fat, r * 1e-3, coarse, sigma_skin_override, str(RESULTS_DIR)) for t_fat in t_fat_list for r in elec_r_list]
    if pipelines > 1 and gpus <= 1:
        # several sweep pipelines on the one GPU (a host thread + context each): one point's meshing / file writing / upload
        # overlaps another's solve; rows come back in sweep order
        return sweep.map_points_pipelined(lambda ctx, state, a: run_case(a[0], a[1], a[2], a[3], a[4], ctx=ctx, results_dir=a[5], quiet=True),
                                          points, device=0, pipelines=pipelines)
    return sweep.map_points(_point, points, gpus=gpus)


def run_sweep_fixed_topology(p, t_fat_list, elec_r_list, coarse=False, sigma_skin_override=None, ctx=None, results_dir=None):
    """The same sweep with the fat thickness as a NODE DISPLACEMENT: per electrode size one mesh (built for the thickest fat
    layer of the sweep, so no layer is stretched thin), one device mesh and one CSR pattern; every thickness moves the nodes
    (``ptfem_mesh_set_coords``), re-assembles and solves.  Rows come back in the reference's order (thickness outer loop)."""
    ctx = ctx or sweep.worker_context()
    results_dir = Path(results_dir) if results_dir else RESULTS_DIR
    results_dir.mkdir(exist_ok=True)
    t_ref = max(t_fat_list)
    rows = {}
    for r_mm in elec_r_list:
        elec_r = r_mm * 1e-3
        ref_dir = results_dir / f"_topology_r{int(r_mm):04d}um"
        mesh_ref, e1, e2, body_info = build_mesh(p, t_ref, elec_r, ref_dir, coarse=coarse)
        dm = ctx.mesh(mesh_ref.nodes, mesh_ref.tets, mesh_ref.region, mesh_ref.tris, mesh_ref.bcid)
        dm.pattern()
        for t_fat in t_fat_list:
            mesh_k = displace_fat_thickness(mesh_ref, p, t_ref, t_fat)
            dm.set_coords(mesh_k.nodes)
            rows[(t_fat, r_mm)] = run_case(p, t_fat, elec_r, coarse, sigma_skin_override, ctx=ctx, results_dir=results_dir,
                                           prebuilt=(mesh_k, e1, e2, body_info), dmesh=dm)
        dm.close()
    return [rows[(t_fat, r_mm)] for t_fat in t_fat_list for r_mm in elec_r_list]


def save_results(all_results, results_dir=None):
    if not all_results:
        return
    results_dir = Path(results_dir) if results_dir else RESULTS_DIR
    keys = list(all_results[0].keys())
    with open(results_dir / "summary.csv", "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=keys)
        w.writeheader()
        w.writerows(all_results)
    print(f"\nSaved → {results_dir / 'summary.csv'}")
    with open(results_dir / "summary.json", "w") as f:
        json.dump(all_results, f, indent=2, default=lambda x: None if isinstance(x, float) and np.isnan(x) else x)
    print(f"Saved → {results_dir / 'summary.json'}")


def main(argv=None):
    ap = argparse.ArgumentParser(description="Ankle layered slab sweep")
    ap.add_argument("--smoke", action="store_true", help="Single coarse case for quick pipeline check")
    ap.add_argument("--gpus", type=int, default=1, help="shard sweep points over this many GPUs")
    ap.add_argument("--pipelines", type=int, default=1, help="sweep pipelines (host thread + GPU context each) on one GPU")
    ap.add_argument("--fixed-topology", action="store_true",
                    help="fat thickness as node displacement on one mesh per electrode size (pattern and device mesh built once) "
                         "instead of re-meshing every point")
    args = ap.parse_args(argv)
    p = load_params()
    pl = _pl(p)
    if args.smoke:
        t_fat_list = [p["layers"]["t_fat"]]
        r_list = [pl.get("electrode_r_mm_list", pl.get("size_list", [10]))[1]]
        print("=== SMOKE TEST (1 coarse case) ===")
    else:
        t_fat_list = p["layers"]["t_fat_sweep"]
        r_list = pl.get("electrode_r_mm_list", pl.get("size_list", [5, 10, 15]))
        print(f"=== FULL SWEEP: {len(t_fat_list)} fat thicknesses × {len(r_list)} electrode sizes = "
              f"{len(t_fat_list)*len(r_list)} cases ===")
    if args.fixed_topology:
        results = run_sweep_fixed_topology(p, t_fat_list, r_list, coarse=args.smoke)
    else:
        results = run_sweep(p, t_fat_list, r_list, coarse=args.smoke, gpus=args.gpus, pipelines=args.pipelines)
    save_results(results)
    print(f"\n  {len(results)} case(s) computed → results/summary.csv, results/summary.json")
    return results


if __name__ == "__main__":
    main()
