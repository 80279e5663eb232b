import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mm3dgs-slam_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built_lib():
    """Path of the in-tree C-ABI library; built on demand where nvcc exists."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gsr_build", os.path.join(ROOT, "mm3dgs-slam_b200", "build.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    if not os.path.exists(m.SO) or (os.path.exists(m.NVCC) and m.needs_build()):
        m.build()
    return m.SO
