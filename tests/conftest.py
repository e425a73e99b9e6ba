import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def engine():
    """One fisr_b200.Engine (C-ABI context) on cuda:0 for the whole GPU session.  No skip-on-missing-library:
    a GPU box without the built .so must fail loudly."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device in this container (GPU tests run under gpurun)")
    import fisr_b200
    eng = fisr_b200.Engine(0)
    yield eng
    eng.close()
