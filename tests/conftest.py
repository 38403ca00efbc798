import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def built_lib():
    """libmrt.so built in-tree (nvcc cross-compiles without a GPU)."""
    from mri_raytracer_b200 import build, _lib
    build.build()
    return _lib.lib()


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible (no CPU fallback exists)")
    from mri_raytracer_b200 import build
    build.build()
    return torch.device("cuda:0")
