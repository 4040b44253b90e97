"""Path bootstrap shared by the driver scripts: makes the repo root importable so that
``import pelvistim_fem_b200`` works when a script is run from its own step directory."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
