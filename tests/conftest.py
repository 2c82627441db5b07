import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


@pytest.fixture(scope="session")
def oracle_available():
    from oracle import binding
    if not binding.available():
        pytest.skip("oracle/_ref/liboracle.so not built (needs /root/reference): run `make -C oracle`")
    return True
