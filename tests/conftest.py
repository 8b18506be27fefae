"""Test configuration: `gpu` marker, import paths, shared helpers.

`-m "not gpu"` : oracle vs golden vectors, host logic, C-ABI load/symbol checks (no device).
`-m gpu`       : CUDA path vs oracle through the C ABI (run on a B200 via gpurun).
Only tests (and smoke / bench's CPU legs) may import `oracle`.
"""
import json
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "wdbx-py_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


try:   # the gates must not depend on a random seed: hypothesis tests replay a fixed set of examples
    import os

    from hypothesis import settings as _hyp_settings

    _hyp_settings.register_profile("gate", derandomize=True)
    _hyp_settings.register_profile("explore", derandomize=False)
    _hyp_settings.load_profile("explore" if os.environ.get("WDBX_MODEL_RANDOM") else "gate")
except ImportError:
    pass


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden():
    return json.loads((ROOT / "tests" / "golden" / "reference_golden.json").read_text())


@pytest.fixture(scope="session")
def built_lib():
    """Build (if stale) and load libwdbx_b200.so."""
    import __graft_entry__ as ge

    ge.build_cuda()
    import wdbx_b200

    return wdbx_b200.load_library()
