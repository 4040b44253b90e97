"""Relative differences between this engine's step03/step04 tables and the reference's committed ones
(tests/golden/*.json).  Usage: python scripts/compare_tables.py <dir with step03_summary.json, step04_summary.json>"""
import json, sys
from pathlib import Path
import numpy as np
d = Path(sys.argv[1] if len(sys.argv) > 1 else "profiles/r01_sweeps")
G = Path("tests/golden")
for step, keys in (("step03", ["elec_area_mesh_cm2", "jn_used", "compliance_V", "total_current_A", "I_return_A", "roi_mean_J", "roi_mean_E",
                              "peak_J_skin_with_elec", "peak_J_skin_no_elec", "efficiency", "flux_err"]),
                   ("step04", ["jn_used_A_m2", "compliance_V", "contact_impedance_ohm", "I_active_A", "I_return_A", "roi_mean_J", "roi_mean_E",
                              "peak_J_skin_with_elec", "peak_J_skin_no_elec", "charge_density_mC_cm2", "efficiency", "flux_err"])):
    ours = json.load(open(d / f"{step}_summary.json")); gold = json.load(open(G / f"{step}_summary.json"))
    assert len(ours) == len(gold)
    print(f"== {step}: {len(ours)} rows; relative difference (ours - reference) / reference, min .. max over the rows")
    for k in keys:
        if k == "flux_err":     # an error measure near zero: absolute values, not ratios
            a = np.array([r[k] for r in ours]); b = np.array([r[k] for r in gold])
            print(f"   {k:28s} ours {a.min():.4f} .. {a.max():.4f}   reference {b.min():.4f} .. {b.max():.4f}   (absolute; rows above 0.03: {(a > 0.03).sum()})")
            continue
        rel = np.array([(a[k] - b[k]) / b[k] for a, b in zip(ours, gold)])
        print(f"   {k:28s} {rel.min():+8.3f} .. {rel.max():+8.3f}   (reference {gold[0][k]:g} .. {gold[-1][k]:g})")
    same = [k for k in gold[0] if all(a[k] == b[k] for a, b in zip(ours, gold))]
    print("   identical columns:", ", ".join(same))
