#!/usr/bin/env python3
"""step01 validation (drop-in for ``step01_box/test_step01_baseline.py``): runs the pipeline if the
VTU is missing, computes the four metrics against the analytic solution V(z) = z/Lz, |J| = sigma dV/Lz
and prints PASS/FAIL with the reference's tolerances (``test_step01_baseline.py:22-25``)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import _common  # noqa: F401,E402
from pelvistim_fem_b200 import pipeline  # noqa: E402

TOL_REL_J, TOL_CV, TOL_R2, TOL_FLUX = 1e-3, 1e-2, 0.9999, 1e-2


def main():
    here = Path.cwd()
    if not (here / "case.sif").exists():
        import setup_case
        setup_case.main(["--mesh"])
    (here / "results").mkdir(exist_ok=True)
    case = pipeline.run_elmer_solver(here)
    m = pipeline.step01_metrics(case)
    checks = [("mean|J| rel err", m["rel_J"], m["rel_J"] < TOL_REL_J), ("CV(|J|)", m["cv_J"], m["cv_J"] < TOL_CV),
              ("R^2 of V(z)", m["r2"], m["r2"] > TOL_R2), ("top/bottom flux mismatch", m["flux_err"], m["flux_err"] < TOL_FLUX)]
    print(f"mean|J| = {m['mean_J']:.6f} A/m2 (analytic {m['J_an']:.6f}), slope = {m['slope']:.4f} V/m, "
          f"Phi range [{m['phi_min']:.4g}, {m['phi_max']:.4g}]")
    ok = True
    for name, val, good in checks:
        print(f"  {'PASS' if good else 'FAIL'}  {name:28s} {val:.3e}")
        ok &= good
    print("OVERALL:", "PASS" if ok else "FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
