import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "shakti-fenics_b200"
for p in (str(ROOT), str(PKG)):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The C-ABI library must exist (built by __graft_entry__.build()); build it if missing."""
    from shakti_b200 import capi
    if not capi.LIB_PATH.exists():
        sys.path.insert(0, str(ROOT))
        import __graft_entry__
        __graft_entry__.build()
    return capi.LIB_PATH
