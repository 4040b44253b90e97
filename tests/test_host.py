"""Host-side logic on CPU: SIF writer/parser against the reference's committed case.sif files, Elmer mesh
and VTU I/O, boundary-id detection, labels, sweep sharding, and the C-ABI library surface."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

import pelvistim_fem_b200  # noqa: F401
from oracle import metrics_oracle as mo
from pelvistim_fem_b200 import elmer_io, engine, meshgen, pipeline, sif, sweep, vtu


# -- SIF -------------------------------------------------------------------------------------------
def test_sif_step01_step02_byte_exact(golden):
    assert sif.serialize(sif.box_case([2], [1])) == (golden / "step01_case.sif").read_text()
    assert sif.serialize(sif.electrode_case(101, 102)) == (golden / "step02_circle_r05mm_case.sif").read_text()


@pytest.mark.parametrize("name,r_mm", [("tfat0003um_r0005um", 5), ("tfat0005um_r0010um", 10), ("tfat0008um_r0015um", 15)])
def test_sif_step03_byte_exact(golden, name, r_mm):
    ref = (golden / f"step03_{name}_case.sif").read_text()
    rep = (golden / f"step03_{name}_bc_debug_report.txt").read_text()
    jn = float(re.search(r"Current Density = (\S+)", ref).group(1))
    area = 5e-3 / jn
    secs, jn_used, _ = sif.layered_case(101, 102, 0.35, 0.04, 0.001, 0.005, elec_r=r_mm * 1e-3, elec_area_mesh=area)
    # the comment carries the mesh area to 4 decimals: take it from the golden report
    a_txt = re.search(r"Mesh area — active electrode : (\S+) cm²", rep).group(1)
    out = sif.serialize(secs)
    assert f"A_mesh={a_txt}cm²" in ref
    assert out == ref


@pytest.mark.parametrize("lvl", ["p01", "p08", "p15"])
def test_sif_step04_byte_exact(golden, lvl):
    import yaml
    ref = (golden / f"step04_{lvl}_case.sif").read_text()
    p = yaml.safe_load((golden / "step04_params.yaml").read_text())
    k = p["pressure_sweep"]["labels"].index(lvl)
    sigma_c = p["pressure_sweep"]["sigma_contact_Spm"][k]
    jn = float(re.search(r"Current Density = (\S+)", ref).group(1))
    secs, _, _ = sif.layered_case(101, 102, 0.35, 0.04, 0.001, sigma_c, elec_r=0.010, elec_area_mesh=5e-3 / jn, dialect="step04")
    assert sif.serialize(secs) == ref


def test_sif_parse_problem(golden):
    pr = sif.problem_from_sif((golden / "step03_tfat0005um_r0010um_case.sif").read_text())
    assert pr.sigma_by_body == {1: 0.35, 2: 0.04, 3: 0.001, 4: 0.005, 5: 0.005}
    assert pr.dirichlet == [(102, 0.0)] and pr.neumann == [(101, 15.97501)] and pr.calc_current
    assert (pr.mesh_db, pr.results_dir, pr.output_name) == ("elmer_mesh", "results", "case")
    p1 = sif.problem_from_sif((golden / "step01_case.sif").read_text())
    assert p1.sigma_by_body == {1: 0.2} and p1.dirichlet == [(2, 1.0), (1, 0.0)] and p1.neumann == []
    multi = sif.serialize(sif.box_case([2, 7], [1]))
    assert "Target Boundaries(2) = 2 7" in multi
    assert sif.problem_from_sif(multi).dirichlet == [(2, 1.0), (7, 1.0), (1, 0.0)]
    with pytest.raises(ValueError):
        sif.problem_from_sif("Header\n  Mesh DB \".\" \"m\"\n")          # missing End
    with pytest.raises(ValueError):
        sif.problem_from_sif("Body 1\n  Target Bodies(1) = 1\n  Material = 3\nEnd\n")   # no conductivity


# -- mesh / VTU I/O ----------------------------------------------------------------------------------
def test_elmer_mesh_roundtrip(tmp_path):
    m = meshgen.synth_slab("XS")
    elmer_io.write_elmer_mesh(tmp_path / "elmer_mesh", m)
    hdr = (tmp_path / "elmer_mesh" / "mesh.header").read_text().split()
    assert [int(hdr[0]), int(hdr[1]), int(hdr[2])] == [m.nn, m.nt, m.nb]
    first = (tmp_path / "elmer_mesh" / "mesh.elements").read_text().splitlines()[0].split()
    assert first[2] == "504" and len(first) == 7                  # find_boundaries.py:31-40
    b = (tmp_path / "elmer_mesh" / "mesh.boundary").read_text().splitlines()[0].split()
    assert b[4] == "303" and len(b) == 8                          # find_boundaries.py:87-90
    r = elmer_io.read_elmer_mesh(tmp_path / "elmer_mesh")
    assert np.array_equal(r.nodes, m.nodes) and np.array_equal(r.tets, m.tets) and np.array_equal(r.region, m.region)
    assert np.array_equal(r.tris, m.tris) and np.array_equal(r.bcid, m.bcid)
    with pytest.raises(FileNotFoundError):
        elmer_io.read_elmer_mesh(tmp_path / "nope")


def test_vtu_roundtrip(tmp_path):
    m = meshgen.box_mesh(nx=3, ny=3, nz=2)
    phi = np.arange(m.nn, dtype=np.float64)
    J = np.random.default_rng(0).standard_normal((m.nn, 3))
    pipeline.write_case_vtu(tmp_path / "case_t0001.vtu", m, phi, J)
    v = vtu.read_vtu(tmp_path / "case_t0001.vtu")
    tets, tris = vtu.split_cells(v)
    assert np.array_equal(v["points"], m.nodes) and np.array_equal(tets, m.tets) and np.array_equal(tris, m.tris)
    assert np.array_equal(v["point_data"]["potential"], phi) and np.array_equal(v["point_data"]["volume current"], J)
    assert list(v["cell_types"][:m.nt]) == [10] * m.nt and list(v["cell_types"][m.nt:]) == [5] * m.nb
    assert v["cell_data"]["GeometryIds"].shape[0] == m.nt + m.nb


# -- pre-solve host logic -----------------------------------------------------------------------------
def test_detect_elec_bc_ids_matches_loop_restatement():
    m = meshgen.layered_slab_mesh(elec_r=0.010, n_muscle=4, n_fat=2, h_bulk=0.006, h_elec=0.003)
    e1, e2 = [0.015, 0.045, 0.0405], [0.065, 0.045, 0.0405]
    got = pipeline.detect_elec_bc_ids(m, e1, e2, 0.0405, 0.0405)
    want = mo.detect_elec_bc_ids_loops(m.nodes, m.tris, m.bcid, e1, e2, 0.0405, 0.0405)
    assert got[:2] == want[:2] == (101, 102)
    assert np.isclose(got[2], want[2], rtol=1e-12) and np.isclose(got[3], want[3], rtol=1e-12)
    assert abs(got[2] - np.pi * 0.010 ** 2) / (np.pi * 0.010 ** 2) < 0.05     # golden: 3.1299 vs 3.1416 cm2
    # swapped positions pick swapped ids
    assert pipeline.detect_elec_bc_ids(m, e2, e1, 0.0405, 0.0405)[:2] == (102, 101)


def test_step01_boundary_classification_and_labels():
    m = meshgen.box_mesh(ids=(2, 1, 3))
    assert pipeline.classify_flat_boundaries(m) == ([2], [1])
    import run_layered_sweep as s3
    assert s3.case_label(0.005, 0.010) == "tfat0005um_r0010um"        # run_layered_sweep.py:1063-1064
    assert s3.case_label(0.003, 0.005) == "tfat0003um_r0005um"
    assert f"{'circle'}_r{int(0.005*1000):02d}mm" == "circle_r05mm"   # run_sweep.py:303


def test_bc_debug_report_byte_exact(golden, tmp_path):
    import yaml
    p = yaml.safe_load((golden / "step03_params.yaml").read_text())
    ref = (golden / "step03_tfat0005um_r0010um_bc_debug_report.txt").read_text()
    jn = float(re.search(r"\(Jn\) : (\S+) A", ref).group(1))
    bi = dict(contact_enabled=True, z_skin_top=0.040, z_elec_top=0.0405, z_e1_skin=0.040, z_e2_skin=0.040,
              z_e1_elec_top=0.0405, z_e2_elec_top=0.0405)
    out = pipeline.save_bc_debug_report(tmp_path, "tfat0005um_r0010um", 101, 102, 5e-3 / jn, 5e-3 / jn, jn, p, bi)
    assert out.read_text() == ref


def test_sweep_assignment():
    assert sweep.assign(9, 4) == [[0, 4, 8], [1, 5], [2, 6], [3, 7]]
    assert sweep.assign(2, 8)[:3] == [[0], [1], []]
    assert sweep.map_points(lambda x: x * x, [1, 2, 3], gpus=1) == [1, 4, 9]


def test_meshers_are_valid():
    for m in (meshgen.box_mesh(jitter=0.3), meshgen.synth_slab("XS"), meshgen.electrode_box_mesh(0.15, 0.15, 0.05, (0.045, 0.075), (0.105, 0.075), 0.01, "square", nz=4)):
        assert (meshgen.tet_volumes(m.nodes, m.tets) > 0).all()
        ext, _ = meshgen.external_faces(m.tets)
        # every external face carries exactly one boundary triangle
        key = lambda t: set(map(tuple, np.sort(t, axis=1).tolist()))
        assert key(ext) <= key(m.tris)
    slab = meshgen.synth_slab("XS")
    assert sorted(np.unique(slab.region).tolist()) == [1, 2, 3, 4, 5] and sorted(np.unique(slab.bcid).tolist()) == [101, 102, 103]
    with pytest.raises(ValueError):
        meshgen.layered_slab_mesh(t_fat=0.0385)


# -- C-ABI surface ---------------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    lib = engine.load_library()
    names = engine.exported_symbols()
    assert len(names) >= 43
    for n in names:
        assert hasattr(lib, n), f"libptfem.so does not export {n}"
    assert lib.ptfem_version() >= 100


def test_no_cpu_fallback_without_device():
    n = ctypes.c_int(-1)
    engine.load_library().ptfem_device_count(ctypes.byref(n))
    if n.value > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(engine.PtfemError) as ei:
        engine.Context(0)
    assert "no CPU fallback" in str(ei.value)


# -- Gmsh .msh reader / ElmerGrid 14 2 ---------------------------------------------------------------------
MSH41 = """$MeshFormat
4.1 0 8
$EndMeshFormat
$PhysicalNames
2
2 101 "active"
3 1 "tissue"
$EndPhysicalNames
$Entities
0 0 2 1
7 0 0 0 1 1 0 1 101 0
8 0 0 0 1 0 1 0 0
1 0 0 0 1 1 1 1 1 2 7 8
$EndEntities
$Nodes
2 5 1 9
2 7 0 3
1
2
3
0 0 0
1 0 0
0 1 0
3 1 0 2
4
9
0 0 1
1 1 1
$EndNodes
$Elements
3 4 1 4
2 7 2 1
1 1 2 3
2 8 2 1
2 1 2 4
3 1 4 2
3 1 2 3 4
4 2 3 4 9
$EndElements
"""


def test_msh41_reader(tmp_path):
    from pelvistim_fem_b200 import gmsh_io
    (tmp_path / "m.msh").write_text(MSH41)
    m = gmsh_io.read_msh(tmp_path / "m.msh")
    assert m.nn == 5 and m.nt == 2 and m.nb == 1             # the triangle on the unnamed surface 8 is dropped
    assert m.bcid.tolist() == [101] and m.region.tolist() == [1, 1]   # named groups keep their ids
    assert np.allclose(m.nodes[4], [1, 1, 1])                 # node tag 9 -> compact index 4
    assert (meshgen.tet_volumes(m.nodes, m.tets) > 0).all()
    assert m.tri_parent.tolist() == [0]


def _write_msh41_the_way_gmsh_lays_it_out(path, m, names):
    """MSH 4.1 ASCII with everything a file written by ``gmsh.write`` after an OpenCASCADE fragment carries and the hand-typed
    sample above does not: point and curve entities (with their own line layouts) ahead of the surfaces, one surface / volume
    entity per physical group plus an untagged surface, a NEGATIVE physical tag (reversed orientation), node blocks classified
    by entity with parametric coordinates on curves and surfaces (extra u / u v columns), node tags numbered entity by entity
    (not in the order the elements use them), and point / line element blocks that a 3-D conversion drops."""
    rng = np.random.default_rng(5)
    bcs, regs = sorted(set(m.bcid.tolist())), sorted(set(m.region.tolist()))
    lo, hi = m.nodes.min(axis=0), m.nodes.max(axis=0)
    bb = " ".join(f"{v:.16g}" for v in (*lo, *hi))
    surf_ent = {b: 10 + k for k, b in enumerate(bcs)}
    vol_ent = {r: 1 + k for k, r in enumerate(regs)}
    L = ["$MeshFormat", "4.1 0 8", "$EndMeshFormat", "$PhysicalNames", str(len(bcs) + len(regs))]
    L += [f'2 {b} "{names.get(b, "b%d" % b)}"' for b in bcs] + [f'3 {r} "body{r}"' for r in regs] + ["$EndPhysicalNames"]
    L += ["$Entities", f"2 1 {len(bcs) + 1} {len(regs)}",
          f"1 {lo[0]:.16g} {lo[1]:.16g} {lo[2]:.16g} 0", f"2 {hi[0]:.16g} {hi[1]:.16g} {hi[2]:.16g} 0",
          f"1 {bb} 0 2 1 -2"]
    for k, b in enumerate(bcs):
        L.append(f"{surf_ent[b]} {bb} 1 {-b if k == 0 else b} 1 1")
    L.append(f"99 {bb} 0 1 1")                                     # a surface without a physical group
    for r in regs:
        L.append(f"{vol_ent[r]} {bb} 1 {r} {len(bcs)} " + " ".join(str(surf_ent[b]) for b in bcs))
    L.append("$EndEntities")
    # classify nodes: first node -> point entity 1, next three -> curve 1, nodes of boundary triangles -> their surface, rest -> volume
    owner = np.full(m.nn, -1)
    kind = {}
    owner[0], kind[0] = 0, (0, 1)
    for i in (1, 2, 3):
        owner[i], kind[i] = 1, (1, 1)
    for t, b in zip(m.tris, m.bcid):
        for i in t:
            if owner[i] < 0:
                owner[i], kind[i] = 2, (2, surf_ent[int(b)])
    first_tet = {}
    for e, t in enumerate(m.tets):
        for i in t:
            first_tet.setdefault(int(i), e)
    for i in range(m.nn):
        if owner[i] < 0:
            owner[i], kind[i] = 3, (3, vol_ent[int(m.region[first_tet[i]])])
    blocks = {}
    for i in range(m.nn):
        blocks.setdefault(kind[i], []).append(i)
    tag = np.zeros(m.nn, dtype=np.int64)
    nxt = 1
    node_lines = []
    for key in sorted(blocks):
        ids = blocks[key]
        dim = key[0]
        par = 1 if dim in (1, 2) else 0
        node_lines.append(f"{dim} {key[1]} {par} {len(ids)}")
        for i in ids:
            tag[i] = nxt
            nxt += 1
        node_lines += [str(tag[i]) for i in ids]
        for i in ids:
            x = " ".join(f"{v:.17g}" for v in m.nodes[i])
            node_lines.append(x + ("" if not par else " " + " ".join(f"{v:.6g}" for v in rng.random(dim))))
    L += ["$Nodes", f"{len(blocks)} {m.nn} 1 {m.nn}"] + node_lines + ["$EndNodes"]
    eb, et = [], 1
    eb.append("0 1 15 1"); eb.append(f"{et} {tag[0]}"); et += 1                               # a point element
    eb.append("1 1 1 2"); eb += [f"{et} {tag[1]} {tag[2]}", f"{et + 1} {tag[2]} {tag[3]}"]; et += 2   # two line elements
    for b in bcs:
        sel = np.nonzero(m.bcid == b)[0]
        eb.append(f"2 {surf_ent[b]} 2 {len(sel)}")
        for k in sel:
            eb.append(f"{et} " + " ".join(str(tag[i]) for i in m.tris[k])); et += 1
    eb.append("2 99 2 1"); eb.append(f"{et} " + " ".join(str(tag[i]) for i in m.tris[0])); et += 1   # triangle on the untagged surface
    for r in regs:
        sel = np.nonzero(m.region == r)[0]
        eb.append(f"3 {vol_ent[r]} 4 {len(sel)}")
        for k in sel:
            eb.append(f"{et} " + " ".join(str(tag[i]) for i in m.tets[k])); et += 1
    nblk = 2 + len(bcs) + 1 + len(regs)
    L += ["$Elements", f"{nblk} {et - 1} 1 {et - 1}"] + eb + ["$EndElements"]
    Path(path).write_text("\n".join(L) + "\n")


def test_msh41_reader_on_a_file_laid_out_as_gmsh_writes_it(tmp_path):
    from pelvistim_fem_b200 import gmsh_io
    m = meshgen.synth_slab("XS")
    _write_msh41_the_way_gmsh_lays_it_out(tmp_path / "mesh.msh", m, {101: "active", 102: "return", 103: "interfaces"})
    r = gmsh_io.read_msh(tmp_path / "mesh.msh")
    assert (r.nn, r.nt, r.nb) == (m.nn, m.nt, m.nb)                # untagged triangle, point and line elements dropped
    assert sorted(set(r.bcid.tolist())) == sorted(set(m.bcid.tolist()))          # named groups keep their ids, sign dropped
    assert np.array_equal(np.sort(r.nodes.view([("", r.nodes.dtype)] * 3), axis=0), np.sort(m.nodes.view([("", m.nodes.dtype)] * 3), axis=0))
    vol = lambda q: {int(k): float(meshgen.tet_volumes(q.nodes, q.tets)[q.region == k].sum()) for k in set(q.region.tolist())}
    va, vb = vol(m), vol(r)
    assert va.keys() == vb.keys() and all(abs(va[k] - vb[k]) <= 1e-12 * abs(va[k]) for k in va)
    area = lambda q: {int(k): float(0.5 * np.linalg.norm(np.cross(q.nodes[q.tris[q.bcid == k]][:, 1] - q.nodes[q.tris[q.bcid == k]][:, 0],
                                                                   q.nodes[q.tris[q.bcid == k]][:, 2] - q.nodes[q.tris[q.bcid == k]][:, 0]), axis=1).sum())
                      for k in set(q.bcid.tolist())}
    aa, ab = area(m), area(r)
    assert aa.keys() == ab.keys() and all(abs(aa[k] - ab[k]) <= 1e-12 * abs(aa[k]) for k in aa)
    assert (meshgen.tet_volumes(r.nodes, r.tets) > 0).all() and (r.tri_parent >= 0).all()
    # and the conversion the reference runs on it gives the same Elmer mesh as from the 2.2 file of the same mesh
    out = gmsh_io.elmergrid_14_2(tmp_path / "mesh.msh", tmp_path / "elmer_mesh")
    assert (tmp_path / "elmer_mesh" / "mesh.header").exists() and out.nt == m.nt


def test_msh22_roundtrip_and_elmergrid_shim(tmp_path):
    import subprocess, sys
    from pathlib import Path
    from pelvistim_fem_b200 import gmsh_io
    m = meshgen.synth_slab("XS")
    names = {(3, 1): "muscle", (3, 2): "fat", (3, 3): "skin", (3, 4): "ca", (3, 5): "cr", (2, 101): "active", (2, 102): "return", (2, 103): "other"}
    gmsh_io.write_msh(tmp_path / "mesh.msh", m, names)
    r = gmsh_io.read_msh(tmp_path / "mesh.msh")
    assert np.array_equal(r.nodes, m.nodes) and np.array_equal(r.tets, m.tets) and np.array_equal(r.tris, m.tris)
    assert np.array_equal(r.region, m.region) and np.array_equal(r.bcid, m.bcid)
    # unnamed groups are renumbered 1..K (box.geo case: boundaries 101/102/103 -> 1/2/3)
    b = meshgen.box_mesh(nx=3, ny=3, nz=2)
    gmsh_io.write_msh(tmp_path / "box.msh", b)
    rb = gmsh_io.read_msh(tmp_path / "box.msh")
    assert sorted(np.unique(rb.bcid).tolist()) == [1, 2, 3] and np.unique(rb.region).tolist() == [1]
    # the shim executable, called exactly as the reference calls ElmerGrid
    shim = Path(__file__).resolve().parent.parent / "drivers" / "bin" / "ElmerGrid"
    pr = subprocess.run([sys.executable, str(shim), "14", "2", "mesh.msh", "-out", "elmer_mesh"], cwd=tmp_path, capture_output=True, text=True)
    assert pr.returncode == 0, pr.stderr
    e = elmer_io.read_elmer_mesh(tmp_path / "elmer_mesh")
    assert np.array_equal(e.tets, m.tets) and np.array_equal(e.bcid, m.bcid) and np.array_equal(e.nodes, m.nodes)
    assert subprocess.run([sys.executable, str(shim), "1", "2", "x.grd"], cwd=tmp_path, capture_output=True).returncode == 1


def test_compression_field_keeps_mesh_valid():
    m = meshgen.synth_slab("S")
    for depth in (0.0005, 0.003):
        n2 = meshgen.compress_under_pads(m.nodes, [(0.015, 0.045), (0.065, 0.045)], 0.010, depth, 0.040)
        assert (meshgen.tet_volumes(n2, m.tets) > 0).all()
        assert np.all(n2[:, 2] <= m.nodes[:, 2] + 1e-15) and np.isclose((m.nodes[:, 2] - n2[:, 2]).max(), depth * 1.0125, rtol=0.02)
        assert np.array_equal(n2[:, :2], m.nodes[:, :2]) and np.all(n2[m.nodes[:, 2] == 0.0, 2] == 0.0)


def test_ctypes_structs_follow_the_header():
    """ptfem_solve_opts / ptfem_solve_stats: field names, order and C types of the ctypes mirrors equal the header's."""
    import ctypes as C
    import re
    from pelvistim_fem_b200 import engine
    from pathlib import Path
    hdr = (Path(__file__).resolve().parents[1] / "include" / "ptfem.h").read_text()
    ctype = {"int32_t": C.c_int32, "double": C.c_double, "int64_t": C.c_int64}
    for name, cls in (("ptfem_solve_opts", engine.SolveOpts), ("ptfem_solve_stats", engine.SolveStats)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = [(m.group(2), ctype[m.group(1)]) for m in re.finditer(r"(int32_t|int64_t|double)\s+(\w+)\s*;", body)]
        assert fields == list(cls._fields_), name


@pytest.mark.parametrize("step,fixture", [("step03_ankle_layers", "step03_params.yaml"), ("step04_pressure", "step04_params.yaml")])
def test_driver_params_carry_the_reference_values(golden, step, fixture):
    # drivers/<step>/params.yaml is written in this repo's own words; keys and values are the reference's
    # (tests/golden/ holds the reference's file as committed there)
    import pathlib
    import yaml
    ours = yaml.safe_load((pathlib.Path(__file__).resolve().parents[1] / "drivers" / step / "params.yaml").read_text())
    assert ours == yaml.safe_load((golden / fixture).read_text())


def test_vtu_bytes_follow_the_vtk_xml_appended_raw_layout(tmp_path):
    # An independent, byte-level reading of the file along the VTK XML file-format rules (VTK File Formats, "XML File Formats":
    # UnstructuredGrid piece; DataArray format="appended" offset=N counts bytes after the '_' that opens <AppendedData
    # encoding="raw">; every block = one header_type (UInt64, little endian) byte count followed by that many raw bytes) - not
    # through vtu.read_vtu.  It is what vtkXMLUnstructuredGridReader does with results/case_t0001.vtu
    # (consumers: step03_ankle_layers/plot_layered_results.py:88-93, smoke_test.py:88-123).
    import re
    import struct
    import xml.etree.ElementTree as ET
    from pelvistim_fem_b200 import meshgen, vtu
    m = meshgen.box_mesh(nx=3, ny=2, nz=2)
    phi = np.linspace(0.0, 1.0, m.nn)
    J = np.stack([phi, 2 * phi, -phi], axis=1)
    gid = np.concatenate([m.region, m.bcid]).astype(np.int32)
    f = tmp_path / "case_t0001.vtu"
    vtu.write_vtu(f, m.nodes, m.tets, m.tris, {"potential": phi, "volume current": J}, {"GeometryIds": gid})
    raw = f.read_bytes()
    cut = raw.index(b'<AppendedData encoding="raw">')
    us = raw.index(b"_", cut) + 1                                   # first byte of the appended data
    tail = raw.rindex(b"</AppendedData>")
    root = ET.fromstring(raw[:cut] + b"</VTKFile>")                 # the XML part alone is well formed
    assert root.tag == "VTKFile" and root.attrib["type"] == "UnstructuredGrid" and root.attrib["byte_order"] == "LittleEndian"
    assert root.attrib["header_type"] == "UInt64"
    piece = root.find("UnstructuredGrid/Piece")
    npts, ncells = int(piece.attrib["NumberOfPoints"]), int(piece.attrib["NumberOfCells"])
    assert npts == m.nn and ncells == m.nt + m.nb
    size = {"Float64": 8, "Float32": 4, "Int32": 4, "Int64": 8, "UInt8": 1}
    np_t = {"Float64": "<f8", "Float32": "<f4", "Int32": "<i4", "Int64": "<i8", "UInt8": "u1"}
    arrays, end = {}, 0
    for sec, ntup in (("PointData", npts), ("CellData", ncells), ("Points", npts), ("Cells", None)):
        for da in piece.find(sec).findall("DataArray"):
            assert da.attrib["format"] == "appended"
            off, t, nc = int(da.attrib["offset"]), da.attrib["type"], int(da.attrib.get("NumberOfComponents", "1"))
            assert off == end, "blocks are laid out back to back in document order"
            (nbytes,) = struct.unpack_from("<Q", raw, us + off)
            if ntup is not None:
                assert nbytes == ntup * nc * size[t], (sec, da.attrib.get("Name"))
            a = np.frombuffer(raw, dtype=np_t[t], count=nbytes // size[t], offset=us + off + 8)
            arrays[(sec, da.attrib.get("Name"))] = a.reshape(-1, nc) if nc > 1 else a
            end = off + 8 + nbytes
    assert us + end <= tail and raw[us + end:tail].strip() == b""    # nothing but white space after the last block
    assert piece.find("Points/DataArray").attrib["NumberOfComponents"] == "3"
    conn, offs, types = arrays[("Cells", "connectivity")], arrays[("Cells", "offsets")], arrays[("Cells", "types")]
    assert types.dtype == np.uint8 and offs.shape[0] == types.shape[0] == ncells
    assert np.all(np.diff(np.concatenate([[0], offs])) == np.where(types == 10, 4, 3)) and offs[-1] == conn.shape[0]
    assert np.all(types[:m.nt] == 10) and np.all(types[m.nt:] == 5)            # tets first, then the boundary triangles
    assert conn.min() >= 0 and conn.max() < npts
    assert np.array_equal(arrays[("Points", None)], m.nodes) and np.array_equal(arrays[("PointData", "potential")], phi)
    assert np.array_equal(arrays[("PointData", "volume current")], J) and np.array_equal(arrays[("CellData", "GeometryIds")], gid)
    assert np.array_equal(conn[:4 * m.nt].reshape(-1, 4), m.tets) and np.array_equal(conn[4 * m.nt:].reshape(-1, 3), m.tris)


# -- sweep pipelines (host threads + contexts on one GPU): host logic with a stand-in for the device context ----------------
class _FakeContext:
    made = []

    def __init__(self, device):
        self.device, self.closed, self.synced = device, False, 0
        _FakeContext.made.append(self)

    def sync(self):
        self.synced += 1

    def close(self):
        self.closed = True


def test_pipeline_pool_order_state_finish_and_errors(monkeypatch):
    import threading
    monkeypatch.setattr(engine, "Context", _FakeContext)
    _FakeContext.made = []
    pool = sweep.PipelinePool(device=3, pipelines=2)
    seen = []

    def fn(ctx, state, pt):
        state["count"] = state.get("count", 0) + 1
        seen.append((state["pipeline"], pt, threading.current_thread().name))
        return (pt, ctx.device, state["pipeline"])

    finished = []
    out = pool.map(fn, range(7), finish=lambda ctx, st: finished.append((st["pipeline"], st["count"])))
    assert [o[0] for o in out] == list(range(7)) and all(o[1] == 3 for o in out)       # results in point order
    assert [o[2] for o in out] == [0, 1, 0, 1, 0, 1, 0]                                # point i -> pipeline i mod P
    assert sorted(finished) == [(0, 4), (1, 3)]                                        # finish once per pipeline, private state
    per = {k: [p for q, p, _ in seen if q == k] for k in (0, 1)}
    assert per[0] == [0, 2, 4, 6] and per[1] == [1, 3, 5]                              # each pipeline keeps its own order
    out2 = pool.map(fn, range(3))                                                      # threads, contexts and state persist
    assert len(_FakeContext.made) == 2 and [o[0] for o in out2] == [0, 1, 2]
    assert sorted(n for _, _, n in seen if n.startswith("ptfem-pipeline"))[0] == "ptfem-pipeline-0"

    def boom(ctx, state, pt):
        if pt == 1:
            raise ValueError("point 1 failed")
        return pt
    with pytest.raises(ValueError, match="point 1 failed"):
        pool.map(boom, range(4))
    assert pool.map(fn, [5])[0][0] == 5                                                # the pool survives a failed map
    pool.close()
    assert all(c.closed for c in _FakeContext.made)
    assert sweep.map_points_pipelined(lambda c, s, p: p * 2, [1, 2, 3], pipelines=2) == [2, 4, 6]


def test_cpulist_parser_and_numa_binding_without_topology():
    assert engine._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert engine._parse_cpulist("") == set()


def test_graded_electrode_box_mesh_is_closed_and_tagged():
    # step02 geometry on the size-field mesher: polygonal patches of the right area, every external face tagged, parents found
    from pelvistim_fem_b200 import sizefield_mesher as sm
    r = 0.010
    m = sm.electrode_box_graded(0.15, 0.15, 0.05, (0.045, 0.075), (0.105, 0.075), r, "circle")
    n = round(2 * np.pi * r / (r / 3.5))
    poly = 0.5 * n * r * r * np.sin(2 * np.pi / n)
    assert abs(m.meta["area_active"] - poly) < 1e-12 and abs(m.meta["area_return"] - poly) < 1e-12
    ext, _ = meshgen.external_faces(m.tets)
    key = lambda a: set(map(tuple, np.sort(a, axis=1).tolist()))
    assert key(ext) == key(m.tris) and (m.tri_parent >= 0).all()
    assert abs(meshgen.tet_volumes(m.nodes, m.tets).sum() - 0.15 * 0.15 * 0.05) < 1e-12
    assert set(np.unique(m.bcid)) == {101, 102, 103}
    sq = sm.electrode_box_graded(0.15, 0.15, 0.05, (0.045, 0.075), (0.105, 0.075), r, "square")
    assert abs(sq.meta["area_active"] - (2 * r) ** 2) < 1e-12                          # square patches are exact
