import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure); compiled on first use."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as _oracle  # noqa: E402

    _oracle.lib()
    return _oracle
