import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'opencl-structure-from-motion_b200')
for p in (ROOT, PKG, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (run on the GPU box with -m gpu)')


@pytest.fixture(scope='session')
def ref():
    """The unmodified reference CPU path (oracle/_ref/libvisoref.so) -- the checker, never the product."""
    import pyref
    return pyref.RefLib()


@pytest.fixture(scope='session')
def ref_nofma():
    import pyref
    return pyref.RefLib('nofma')


@pytest.fixture(scope='session')
def ctx():
    import visocu_py
    c = visocu_py.Context(0)
    yield c
    c.close()
