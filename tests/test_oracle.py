"""The CPU oracle against the reference's known answers: the analytic step01 case
(exact for P1), physical laws (KCL, series resistance of the contact layer as in step04's table), the
golden step03/step04 tables to discretisation accuracy, and the C oracle against the numpy oracle."""
import json

import numpy as np
import pytest
import scipy.sparse as sp

import pelvistim_fem_b200  # noqa: F401
from conftest import SIGMA5
from oracle import c_oracle as co
from oracle import fem_oracle as fo
from oracle import metrics_oracle as mo
from pelvistim_fem_b200 import meshgen


@pytest.mark.parametrize("jitter", [0.0, 0.3])
def test_step01_analytic_exact(jitter):
    # test_step01_baseline.py:59-104: phi = z/Lz, J = (0,0,-sigma/Lz) at every node; published:
    # mean|J| 10.000000, CV 8.5e-16, R^2 1.0000000, slope 50.0000 (step01_summary.png)
    m = meshgen.box_mesh(0.04, 0.04, 0.02, 10, 10, 5, jitter=jitter, seed=3, ids=(2, 1, 3))
    res = fo.solve_case(m, {1: 0.2}, [(2, 1.0), (1, 0.0)], [], recover="l2")
    assert np.abs(res["phi"] - m.nodes[:, 2] / 0.02).max() < 1e-12
    assert np.abs(res["J"] - np.array([0, 0, -10.0])).max() < 1e-9
    met = mo.step01_metrics(m.nodes, res["phi"], res["J"])
    assert met["rel_J"] < 1e-3 and met["cv_J"] < 1e-2 and met["r2"] > 0.9999 and met["flux_err"] < 1e-2   # :22-25
    assert abs(met["slope"] - 50.0) < 1e-9 and abs(met["mean_J"] - 10.0) < 1e-9


def test_pattern_and_matrix_properties():
    m = meshgen.synth_slab("XS")
    rowptr, col = fo.csr_pattern(m.nn, m.tets)
    K = fo.assemble_stiffness(m.nodes, m.tets, m.region, SIGMA5)
    assert np.array_equal(K.indptr, rowptr) and np.array_equal(K.indices, col)
    for i in range(m.nn):                                     # sorted, diagonal present
        c = col[rowptr[i]:rowptr[i + 1]]
        assert np.all(np.diff(c) > 0) and i in c
    assert abs(K - K.T).max() < 1e-15 * abs(K).max() * 10
    assert np.abs(K @ np.ones(m.nn)).max() < 1e-12 * abs(K).max()   # constants are in the null space
    M = fo.assemble_mass(m.nodes, m.tets)
    vol, _ = fo.tet_geometry(m.nodes, m.tets)
    assert abs(M.sum() - vol.sum()) < 1e-15 * m.nn


def test_kcl_and_series_resistance_law():
    # weak-form reaction at the Dirichlet electrode = injected current (exact KCL); and the step04 law
    # compliance(sigma_c) ~ c0 + 2 t_c I / (sigma_c A) extracted from step04 summary.csv
    m = meshgen.synth_slab("XS")
    areas = meshgen.tri_areas(m.nodes, m.tris)
    A = areas[m.bcid == 101].sum()
    I = 5e-3
    comp = []
    for sc in (5e-5, 5e-3):
        sig = dict(SIGMA5)
        sig[4] = sig[5] = sc
        r = fo.solve_case(m, sig, [(102, 0.0)], [(101, I / A)], recover=None)
        react = (r["K_raw"] @ r["phi"] - r["b_neumann"])
        nodes102 = np.unique(m.tris[m.bcid == 102])
        assert abs(react[nodes102].sum() + I) < 1e-9 * I * 100
        top = np.unique(m.tris[m.bcid == 101])
        comp.append(r["phi"][top].mean())
    slope = (comp[0] - comp[1]) / (1 / 5e-5 - 1 / 5e-3)
    assert abs(slope - 2 * 0.0005 * I / A) / (2 * 0.0005 * I / A) < 0.05


def test_c_oracle_matches_numpy_oracle():
    m = meshgen.synth_slab("XS")
    ref = fo.solve_case(m, SIGMA5, [(102, 0.0)], [(101, 15.975)], recover=None)
    cs = co.CSystem(m, SIGMA5, [(102, 0.0)], [(101, 15.975)])
    rp, col = fo.csr_pattern(m.nn, m.tets)
    assert np.array_equal(cs.rowptr, rp) and np.array_equal(cs.col, col)
    assert np.abs(cs.val_raw - ref["K_raw"].data).max() <= 4e-15 * np.abs(ref["K_raw"].data).max()
    Kc = sp.csr_matrix((cs.val, cs.col, cs.rowptr), shape=(m.nn, m.nn))
    assert abs(Kc - ref["K"]).max() <= 1e-14 * abs(ref["K"]).max()
    assert np.abs(cs.b - ref["b"]).max() <= 1e-15
    x, it, rel = cs.pcg(1e-12)
    assert rel <= 1e-12 and np.abs(x - ref["phi"]).max() <= 1e-9 * np.abs(ref["phi"]).max()
    v = np.random.default_rng(0).standard_normal(m.nn)
    assert np.abs(cs.spmv(v) - ref["K"] @ v).max() <= 1e-13 * np.abs(ref["K"] @ v).max()


def test_numpy_pcg_matches_direct():
    m = meshgen.synth_slab("XS")
    ref = fo.solve_case(m, SIGMA5, [(102, 0.0)], [(101, 15.975)], recover=None)
    x, it = fo.jacobi_pcg(ref["K"], ref["b"], rtol=1e-12)
    assert np.abs(x - ref["phi"]).max() < 1e-9 * np.abs(ref["phi"]).max()


def test_recovery_variants_on_linear_field():
    # every recovery must reproduce a constant J exactly
    m = meshgen.box_mesh(0.04, 0.04, 0.02, 6, 6, 4, jitter=0.25, seed=1)
    phi = 3.0 * m.nodes[:, 0] - 2.0 * m.nodes[:, 1] + 5.0 * m.nodes[:, 2]
    for method in ("l2", "lumped", "average"):
        J = fo.recover_nodal_current(m.nodes, m.tets, m.region, {1: 0.5}, phi, method)
        assert np.abs(J - (-0.5) * np.array([3.0, -2.0, 5.0])).max() < 1e-11


def test_metrics_hand_cases():
    # one tet + one boundary triangle: VTK-style cell averaging and centre gradients by hand
    pts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.0]])
    tets = np.array([[0, 1, 2, 3]], dtype=np.int32)
    tris = np.array([[0, 1, 2]], dtype=np.int32)
    phi = np.array([0.0, 1.0, 2.0, 3.0])
    J = np.tile(np.array([0.0, 0.0, -2.0]), (4, 1))
    Jm, Em, cen = mo.cell_fields(pts, tets, tris, phi, J)
    assert np.allclose(Jm, 2.0) and np.allclose(cen[0], [0.25, 0.25, 0.25]) and np.allclose(cen[1], [1 / 3, 1 / 3, 0])
    # smoothed point values: nodes 0..2 belong to tet (mean 1.5) and tri (mean 1.0) -> 1.25 ; node 3 -> 1.5
    ps = np.array([1.25, 1.25, 1.25, 1.5])
    assert np.isclose(Em[0], abs(ps[3] - ps[0])) and np.isclose(Em[1], 0.0)
    Ia, Ir, ferr, Ias, Irs = mo.injected_current(pts, tris, J, (1 / 3, 1 / 3), (1 / 3, 1 / 3), 1.0, 0.0, 0.0)
    assert np.isclose(Ias, -2.0 * 0.5)


def test_layered_oracle_vs_golden_table(golden):
    # step03 summary row (t_fat 5 mm, r 10 mm) on the size-field mesh: the pad rim is the polygon Gmsh's 1-D mesh of the circle
    # gives, so the mesh area and the imposed current density equal the reference's to every printed digit; the potential-
    # driven columns agree to discretisation accuracy, the pad-rim peak - set by the element layout at the rim - to a few %
    import run_layered_sweep as s3
    import tempfile
    from pathlib import Path
    from pelvistim_fem_b200 import pipeline, sif
    p = s3.load_params()
    with tempfile.TemporaryDirectory() as d:
        mesh, e1, e2, bi = s3.build_mesh(p, 0.005, 0.010, Path(d) / "c", coarse=False)
        e1id, e2id, Aa, Ar = pipeline.detect_elec_bc_ids(mesh, e1, e2, e1[2], e2[2])
        jn = s3.write_sif(Path(d) / "c", e1id, e2id, p, 0.010, bi, elec_area_mesh=Aa)
        text = (Path(d) / "c" / "case.sif").read_text()
        prob = sif.problem_from_sif(text)
    assert (e1id, e2id) == (101, 102)
    # the whole case.sif - Current Density line with the mesh area in its comment included - is the reference's file
    assert text == (golden / "step03_tfat0005um_r0010um_case.sif").read_text()
    ref = fo.solve_case(mesh, prob.sigma_by_body, prob.dirichlet, prob.neumann, recover="lumped")
    row = mo.layered_row(mesh.nodes, mesh.tets, mesh.tris, ref["phi"], ref["J"], p, 0.005, 0.010, e1, e2, bi, jn_used=jn,
                         elec_area_mesh=Aa, return_area_mesh=Ar, e1_id=e1id, e2_id=e2id)
    gold = [r for r in json.load(open(golden / "step03_summary.json")) if r["t_fat_mm"] == 5.0 and r["elec_r_mm"] == 10.0][0]
    assert list(row.keys()) == list(gold.keys())                       # 36 columns, same order
    for k, tol in (("compliance_V", 0.012), ("roi_mean_J", 0.04), ("peak_J_skin_with_elec", 0.05), ("roi_mean_E", 0.08),
                   ("total_current_A", 0.002), ("I_return_A", 0.02)):
        assert abs(row[k] - gold[k]) / abs(gold[k]) < tol, (k, row[k], gold[k])
    for k in ("elec_shape", "contact_enabled", "control_mode", "roi_layer", "roi_center_z_mm", "dist_fat_muscle_mm",
              "active_boundary_id_used", "return_boundary_id_used", "elec_area_cm2", "t_fat_mm", "elec_r_mm",
              "elec_area_mesh_cm2", "return_area_mesh_cm2", "jn_used"):
        assert row[k] == gold[k], k
    assert row["flux_err"] < 0.03                                      # reference 0.0088
    # weak-form KCL is exact even though the nodal-J pad integral is not (run_layered_sweep.py README note)
    react = ref["K_raw"] @ ref["phi"] - ref["b_neumann"]
    I_in = prob.neumann[0][1] * Aa            # the SIF holds Jn with 7 significant digits
    assert abs(react[np.unique(mesh.tris[mesh.bcid == 102])].sum() + I_in) < 1e-12


def test_graded_mesher_reproduces_reference_mesh_areas(golden):
    # the polygon areas of the three pad sizes (tangent pads included) equal the reference's mesh areas, 4 printed digits
    from pelvistim_fem_b200 import sizefield_mesher as sm, meshgen as mg
    rows = json.load(open(golden / "step03_summary.json"))
    for r_mm in (5.0, 10.0, 15.0):
        gold = [r for r in rows if r["elec_r_mm"] == r_mm][0]
        m = sm.layered_slab_graded(elec_r=r_mm * 1e-3, t_fat=0.005)
        assert round(m.meta["area_active"] * 1e4, 4) == gold["elec_area_mesh_cm2"]
        assert round(m.meta["area_return"] * 1e4, 4) == gold["return_area_mesh_cm2"]
        assert (m.tri_parent >= 0).all() and (mg.tet_volumes(m.nodes, m.tets) > 0).all()
        ext, _ = mg.external_faces(m.tets)                 # closed surface: every external face is a tagged triangle
        key = lambda a: set(map(tuple, np.sort(a, axis=1).tolist()))
        assert key(ext) <= key(m.tris)
        vol = 0.08 * 0.06 * 0.04 + (m.meta["area_active"] + m.meta["area_return"]) * 0.0005
        assert abs(mg.tet_volumes(m.nodes, m.tets).sum() - vol) < 1e-12 * vol * 1e3
        assert m.meta["triangulation"]["mean_quality"] > 0.95


def test_graded_mesher_matches_the_reference_cell_density_in_the_roi(golden):
    # the reference's tables hold ONE mesh-density figure per row: roi_n_cells, the VTU cells (tets and interface triangles) whose
    # centroid lies in the 5 mm ROI sphere 10 mm under the active pad.  roi_mean_E is a COUNT-weighted mean over those cells and
    # the field is ~9x larger in fat than in muscle, so the column follows the cell density on either side of the fat / muscle
    # interface, not only the solution.  The level spacing of the extruded mesh is calibrated on this figure (muscle: 1.5 x the
    # size field, times (lc(d)/lc_elec)^2 where the reference's 3-D field has grown past the pad size; fat: 1.25 x), NOT on the
    # field columns: every sweep point lands within -10 % .. +25 % of the reference's count (+20 .. +60 % before)
    import run_layered_sweep as s3
    import tempfile
    from pathlib import Path
    p = s3.load_params()
    rows = json.load(open(golden / "step03_summary.json"))
    ratios = []
    with tempfile.TemporaryDirectory() as d:
        for g in rows:
            t_fat, elec_r = g["t_fat_mm"] * 1e-3, g["elec_r_mm"] * 1e-3
            mesh, e1, e2, bi = s3.build_mesh(p, t_fat, elec_r, Path(d) / f"c{len(ratios)}")
            cen = np.array([e1[0], e1[1], bi["z_skin_top"] - p["roi"]["z_target"]])
            c = np.concatenate([mesh.nodes[mesh.tets].mean(axis=1), mesh.nodes[mesh.tris].mean(axis=1)])
            n = int((np.linalg.norm(c - cen, axis=1) < p["roi"]["roi_radius"]).sum())
            ratios.append(n / g["roi_n_cells"])
            assert 0.90 < ratios[-1] < 1.25, (g["t_fat_mm"], g["elec_r_mm"], n, g["roi_n_cells"])
    assert abs(np.mean(ratios) - 1.0) < 0.08


def test_step04_series_law_from_golden(golden):
    # the law extracted from the reference's own table: compliance_V(sigma_c) - compliance_V(p15) ~ 2 t_c I/(sigma_c A)
    rows = json.load(open(golden / "step04_summary.json"))
    A, I, tc = 3.1299e-4, 5e-3, 0.5e-3
    base = rows[-1]["compliance_V"] - 2 * tc * I / (rows[-1]["sigma_contact_Spm"] * A)
    for r in rows:
        pred = base + 2 * tc * I / (r["sigma_contact_Spm"] * A)
        assert -0.04 < (pred - r["compliance_V"]) / r["compliance_V"] < 0.005


# -- unstructured (Delaunay) meshes: irregular connectivity like a Gmsh mesh ---------------------------------------
from hypothesis import given, settings, strategies as st


@settings(max_examples=6, deadline=None)
@given(seed=st.integers(0, 10_000), npts=st.integers(150, 600))
def test_property_unstructured_meshes(seed, npts):
    m = meshgen.delaunay_box_mesh(npts, seed=seed)
    vol = meshgen.tet_volumes(m.nodes, m.tets)
    assert (vol > 0).all() and abs(vol.sum() - 0.04 * 0.03 * 0.02) < 1e-4 * 2.4e-5   # box filled (dropped slivers have ~0 volume)
    rowptr, col = fo.csr_pattern(m.nn, m.tets)
    cs = co.CSystem(m, {1: 0.2}, [(101, 1.0), (102, 0.0)], [])
    assert np.array_equal(cs.rowptr, rowptr) and np.array_equal(cs.col, col)         # C and numpy oracles agree bit for bit
    ref = fo.solve_case(m, {1: 0.2}, [(101, 1.0), (102, 0.0)], [], recover="l2")
    assert np.abs(cs.val_raw - ref["K_raw"].data).max() <= 1e-13 * np.abs(ref["K_raw"].data).max()
    # the analytic solution is exact on ANY tet mesh (test_step01_baseline.py: V = z/Lz, J = -sigma/Lz)
    assert np.abs(ref["phi"] - m.nodes[:, 2] / 0.02).max() < 1e-11
    assert np.abs(ref["J"] - np.array([0.0, 0.0, -10.0])).max() < 1e-8
    x, it, rel = cs.pcg(1e-13)
    assert np.abs(x - ref["phi"]).max() < 1e-9


def test_coarse_grid_preconditioner_restatement():
    """oracle/coarse_oracle.py: same answer as the direct solve, several times fewer iterations than Jacobi, and the
    multilevel variant no worse than the two-level one (the CUDA solver's iteration counts are compared with these)."""
    from oracle import coarse_oracle as cz
    m = meshgen.synth_slab("S", interfaces_as_103=False)
    K_raw = fo.assemble_stiffness(m.nodes, m.tets, m.region, SIGMA5)
    is_dir, val = fo.dirichlet_nodes(m.tris, m.bcid, [(102, 0.0)], m.nn)
    b = fo.neumann_rhs(m.nodes, m.tris, m.bcid, [(101, 10.0)])
    K, b = fo.apply_dirichlet_symmetric(K_raw, b, is_dir, val)
    K = K.tocsr()
    ref = fo.solve_direct(K, b)
    dinv = 1.0 / K.diagonal()
    x_j, it_j = cz.pcg(K, b, lambda r: dinv * r)
    counts = {}
    for levels in (0, 1):
        M = cz.CoarsePreconditioner(K, m.nodes, is_dir, coarse_nodes=300, extra_levels=levels)
        assert 200 <= M.coarse_unknowns <= 450 and M.nlev == levels + 1
        x, it = cz.pcg(K, b, M.apply)
        assert np.abs(x - ref).max() <= 1e-8 * np.abs(ref).max()
        counts[levels] = it
    assert np.abs(x_j - ref).max() <= 1e-8 * np.abs(ref).max()
    assert counts[0] * 3 < it_j and counts[1] <= counts[0]
    # grid choice follows coarse.cu: >= 2 cells per axis, about the requested node count
    n = cz.choose_grid(np.zeros(3), np.array([0.08, 0.06, 0.0405]), 2000.0)
    assert (n >= 2).all() and 1500 <= np.prod(n + 1) <= 2600


def test_c_oracle_coarse_grid_pcg_and_recovery_match_numpy_oracle():
    # oracle/fem_c.c oc_pcg_coarse / oc_recover_lumped (what bench.py's CPU legs and the size-M/L GPU parity tests use)
    # against the numpy oracle: direct solve, coarse_oracle.py iteration counts, lumped recovery
    from oracle import coarse_oracle as cor
    co.use_all_cores()
    m = meshgen.synth_slab("S")
    ref = fo.solve_case(m, SIGMA5, [(102, 0.0)], [(101, 15.975)], recover="lumped")
    cs = co.CSystem(m, SIGMA5, [(102, 0.0)], [(101, 15.975)])
    K = sp.csr_matrix((cs.val, cs.col, cs.rowptr), shape=(cs.nn, cs.nn))
    for nodes, levels in ((300, 0), (120, 1)):
        c = cs.coarse_setup(coarse_nodes=nodes, extra_levels=levels)
        x, it, rel = cs.pcg_coarse(rtol=1e-12)
        P = cor.CoarsePreconditioner(K, m.nodes, cs.isdir.astype(bool), coarse_nodes=nodes, extra_levels=levels)
        xn, itn = cor.pcg(K, cs.b, P.apply, rtol=1e-12)
        assert c["nlev"] == P.nlev == levels + 1 and c["coarse_unknowns"] == P.coarse_unknowns
        assert abs(it - itn) <= 1 and rel <= 1e-12
        assert np.abs(x - ref["phi"]).max() < 1e-9 * np.abs(ref["phi"]).max()
    J = cs.recover_lumped(ref["phi"])
    assert np.abs(J - ref["J"]).max() < 1e-12 * np.abs(ref["J"]).max()
    # thread count is explicit: an exported OMP_NUM_THREADS=1 (torch.distributed.run does that) must not stick
    co.set_threads(1)
    assert co.threads() == 1
    assert co.use_all_cores() == co.host_cores() == co.threads()


def _bone_layer_case():
    """Layered slab whose bone block is a FULL layer inside the muscle, 1 V across top and bottom faces: a 1-D series
    circuit with a known answer that linear tets reproduce exactly (interfaces lie on mesh planes)."""
    m = meshgen.layered_slab_mesh(bone=dict(x=(0, 0.08), y=(0, 0.06), z=(0.010, 0.020)), contact_enabled=False,
                                  h_bulk=0.006, h_elec=0.004, n_muscle=10)
    zt = m.nodes[m.tris][:, :, 2]
    bc = np.where(np.all(np.abs(zt - 0.040) < 1e-12, axis=1), 201, np.where(np.all(np.abs(zt) < 1e-12, axis=1), 202, 103)).astype(np.int32)
    m = meshgen.TetMesh(m.nodes, m.tets, m.region, m.tris, bc, meta=m.meta)
    sig = {1: 0.35, 2: 0.04, 3: 0.001, 6: 0.02}
    zb0, zb1 = m.meta["bone"]["z"]
    z_mf, z_fs, Lz = 0.040 - 0.0015 - 0.005, 0.040 - 0.0015, 0.040
    layers = [(0.0, zb0, sig[1]), (zb0, zb1, sig[6]), (zb1, z_mf, sig[1]), (z_mf, z_fs, sig[2]), (z_fs, Lz, sig[3])]
    R = sum((b - a) / s for a, b, s in layers)
    Jz = -1.0 / R                                  # 1 V on top, 0 V at the bottom: current flows down

    def phi_exact(z):
        out = np.zeros_like(z)
        acc = 0.0
        for a, b, s in layers:
            out = np.where(z > a, acc + (np.minimum(z, b) - a) / s / R, out)
            acc += (b - a) / s / R
        return out
    return m, sig, phi_exact, Jz


def test_bone_layer_series_circuit_exact():
    # the bone body (region 6): analytic two-(five-)material known answer, exact for P1
    m, sig, phi_exact, Jz = _bone_layer_case()
    assert 6 in np.unique(m.region)
    r = fo.solve_case(m, sig, [(201, 1.0), (202, 0.0)], [], recover="lumped")
    assert np.abs(r["phi"] - phi_exact(m.nodes[:, 2])).max() < 1e-11
    assert np.abs(r["J"][:, 2] - Jz).max() < 1e-9 * abs(Jz) and np.abs(r["J"][:, :2]).max() < 1e-9 * abs(Jz)
