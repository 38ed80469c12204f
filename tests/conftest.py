import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope='session')
def product():
    """The product package as a namespace with the reference's module names."""
    import types
    import ies_b200
    return types.SimpleNamespace(space=ies_b200.space, source=ies_b200.source,
                                 structure=ies_b200.structure, collector=ies_b200.collector)
