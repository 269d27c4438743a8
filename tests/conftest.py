import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def mb():
    import muscade_b200
    return muscade_b200


@pytest.fixture(scope="session")
def engine_factory(mb):
    made = []

    def make():
        e = mb.Engine(0)
        made.append(e)
        return e
    yield make
    for e in made:
        e.close()
