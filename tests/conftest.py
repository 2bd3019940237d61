import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def host():
    """brille's own module built from the reference sources (oracle/_ref); skip when it was never built."""
    from oracle import ref

    if not ref.available():
        pytest.skip("reference build (oracle/_ref) not available")
    return ref.host()


@pytest.fixture(scope="session")
def probe(host):
    from oracle import ref

    return ref.probe()


@pytest.fixture(scope="session")
def bridge():
    try:
        from brille_b200 import _bridge
    except ImportError:
        pytest.skip("brille_b200._bridge not built")
    return _bridge
