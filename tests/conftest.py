import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def built_library():
    """The in-tree CUDA library; built here if missing (nvcc cross-compiles without a GPU)."""
    from robustgrape_b200 import build
    return build.build_cuda()


@pytest.fixture(scope="session")
def gpu_ctx(built_library):
    from robustgrape_b200._lib import default_context
    return default_context(0)
