import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden():
    import torch
    return torch.load(ROOT / "tests" / "golden" / "reference_golden.pt", weights_only=False)


@pytest.fixture(scope="session")
def ug():
    """The C-ABI library through the host-side ops layer; fails loudly (no fallback) when it is missing."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from unigen_b200 import ops
    return ops
