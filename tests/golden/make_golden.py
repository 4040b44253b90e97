"""Copy the reference's committed golden artefacts into tests/golden/ (run in the
build container only; /root/reference does not exist on the GPU box).

Golden data = files the reference itself produced and committed:
generated case.sif files, bc_debug_report.txt, summary.csv/json tables
(SURVEY.md section 8c).  No reference *source* is copied.
"""
import shutil
from pathlib import Path

REF = Path("/root/reference")
OUT = Path(__file__).parent

ITEMS = [
    ("step01_box/case.sif", "step01_case.sif"),
    ("step02_electrodes/results/circle_r05mm/case.sif", "step02_circle_r05mm_case.sif"),
    ("step03_ankle_layers/results/summary.csv", "step03_summary.csv"),
    ("step03_ankle_layers/results/summary.json", "step03_summary.json"),
    ("step04_pressure/results/summary.csv", "step04_summary.csv"),
    ("step04_pressure/results/summary.json", "step04_summary.json"),
    ("step03_ankle_layers/params.yaml", "step03_params.yaml"),
    ("step04_pressure/params.yaml", "step04_params.yaml"),
]
for case in ("tfat0003um_r0005um", "tfat0005um_r0010um", "tfat0008um_r0015um"):
    ITEMS.append((f"step03_ankle_layers/results/{case}/case.sif", f"step03_{case}_case.sif"))
    ITEMS.append((f"step03_ankle_layers/results/{case}/bc_debug_report.txt", f"step03_{case}_bc_debug_report.txt"))
for lvl in ("p01", "p08", "p15"):
    ITEMS.append((f"step04_pressure/results/{lvl}/case.sif", f"step04_{lvl}_case.sif"))

# step02_png_titles.json is transcribed by hand from step02_electrodes/results/sweep_J_maps.png

if __name__ == "__main__":
    for src, dst in ITEMS:
        shutil.copyfile(REF / src, OUT / dst)
        print(f"{src} -> tests/golden/{dst}")
