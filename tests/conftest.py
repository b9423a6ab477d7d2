import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "synthetic-sleep-eeg-signal-generation-using-latent-diffusion-models_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

REFERENCE = "/root/reference"
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference (only present in the build container)")


def pytest_collection_modifyitems(config, items):
    has_ref = os.path.isdir(os.path.join(REFERENCE, "src", "models"))
    skip_ref = pytest.mark.skip(reason="/root/reference not present")
    for item in items:
        if "reference" in item.keywords and not has_ref:
            item.add_marker(skip_ref)


@pytest.fixture(scope="session")
def built_lib():
    """libeegldm.so, built on demand (nvcc cross-compiles without a GPU)."""
    import build as _build  # <package>/build.py
    path = _build.build()
    from eegldm import _lib
    return _lib.lib()


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but torch.cuda.is_available() is False (eegldm has no CPU fallback)")
    return torch.device("cuda", 0)
