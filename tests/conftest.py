import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def cref():
    from oracle import cref as c

    c.lib()
    return c


@pytest.fixture(scope="session")
def eng():
    """The product: ark_blst_b200 over libb200msm.so. Import fails loudly if the .so is absent."""
    import ark_blst_b200 as e

    return e
