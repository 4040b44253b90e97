import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "drivers", ROOT / "drivers" / "step03_ankle_layers", ROOT / "drivers" / "step04_pressure",
          ROOT / "drivers" / "step02_electrodes", ROOT / "drivers" / "step01_box"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return GOLDEN


@pytest.fixture(scope="session")
def gpu_ctx():
    """One libptfem context on cuda:0 for the whole GPU test session (fails loudly without a device)."""
    import pelvistim_fem_b200  # noqa: F401
    from pelvistim_fem_b200 import engine
    ctx = engine.Context(0)
    yield ctx
    ctx.close()


SIGMA5 = {1: 0.35, 2: 0.04, 3: 0.001, 4: 0.005, 5: 0.005}
