import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle
    oracle.build()
    return oracle.restatement()


@pytest.fixture(scope="session")
def engine():
    from rabbitsalign_b200 import ExtensionEngine
    e = ExtensionEngine(device=0)
    yield e
    e.close()


@pytest.fixture(scope="session")
def engine_exact():
    from rabbitsalign_b200 import ExtensionEngine
    e = ExtensionEngine(device=0, exact_only=True)
    yield e
    e.close()
