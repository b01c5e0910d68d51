import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pk():
    import pkb200

    return pkb200.pk


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle_py

    if not os.path.exists(os.path.join(oracle_py.HERE, "liboracle.so")):
        oracle_py.build(ref=False)
    return oracle_py
