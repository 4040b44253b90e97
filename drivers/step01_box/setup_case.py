#!/usr/bin/env python3
"""step01: write ``case.sif`` for the mesh in ``./elmer_mesh`` (drop-in for the reference's
``step01_box/setup_case.py``: run in the case directory, reads ``elmer_mesh/mesh.nodes`` and
``elmer_mesh/mesh.boundary``, detects the boundary ids that are flat at the global zmax / zmin
(``setup_case.py:107-118``) and writes the two-electrode Dirichlet problem of ``case.sif``).

If ``elmer_mesh/`` does not exist and ``--mesh`` is given, the built-in box mesher stands in for
``gmsh -3 box.geo`` + ``ElmerGrid 14 2`` (neither exists where this engine runs)."""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import _common  # noqa: F401,E402
from pelvistim_fem_b200 import elmer_io, meshgen, pipeline, sif  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--mesh", action="store_true", help="generate elmer_mesh/ with the built-in box mesher (box.geo geometry)")
    ap.add_argument("--jitter", type=float, default=0.25, help="interior node jitter of the built-in mesher (fraction of h)")
    args = ap.parse_args(argv)
    meshdir = Path("elmer_mesh")
    if args.mesh and not (meshdir / "mesh.nodes").exists():
        # box.geo:4-21: 0.04 x 0.04 x 0.02 m, lc = 4 mm; ElmerGrid renumbers the unnamed groups to 1..3
        m = meshgen.box_mesh(0.04, 0.04, 0.02, 10, 10, 5, jitter=args.jitter, ids=(2, 1, 3))
        elmer_io.write_elmer_mesh(meshdir, m)
        print(f"built-in mesher: {m.nn} nodes, {m.nt} tets -> {meshdir}/")
    if not (meshdir / "mesh.nodes").exists() or not (meshdir / "mesh.boundary").exists():
        raise SystemExit("ERROR: run this in the folder that contains elmer_mesh/mesh.nodes and elmer_mesh/mesh.boundary")
    mesh = elmer_io.read_elmer_mesh(meshdir)
    z = mesh.nodes[:, 2]
    print(f"Global zmin={z.min():.6g}, zmax={z.max():.6g}, Lz={z.max()-z.min():.6g}")
    top_ids, bot_ids = pipeline.classify_flat_boundaries(mesh)
    print("\nDetected TOP boundary IDs:", top_ids)
    print("Detected BOTTOM boundary IDs:", bot_ids)
    if not top_ids or not bot_ids:
        raise SystemExit("ERROR: Could not confidently detect top/bottom boundary IDs. See summary above.")
    Path("case.sif").write_text(sif.serialize(sif.box_case(top_ids, bot_ids)))
    print("\nWrote case.sif with:")
    print("  TOP    ->", sif.format_target(top_ids))
    print("  BOTTOM ->", sif.format_target(bot_ids))


if __name__ == "__main__":
    main()
