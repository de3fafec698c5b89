import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def built():
    """The C-ABI library and the oracle must exist (built by __graft_entry__.build())."""
    import importlib
    T = importlib.import_module("toy-spice_b200")
    if not os.path.exists(T.lib_path()):
        import __graft_entry__ as g
        g.build()
    T.lib()
    from oracle import oracle as O
    O.lib()
    return T


@pytest.fixture(scope="session")
def ctx(built):
    return built.Context(0)
