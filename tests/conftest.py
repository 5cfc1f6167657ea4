import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def svb():
    import stark_verifier_b200 as m
    if not os.path.exists(m.lib_path()):
        import __graft_entry__ as g
        g.build()
    return m


@pytest.fixture(scope="session")
def orc():
    from oracle import binding
    binding.build()
    return binding


@pytest.fixture(scope="session")
def ctx(svb):
    c = svb.Context(0)
    yield c
    c.close()
