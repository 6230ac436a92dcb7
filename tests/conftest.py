import os, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FIXTURES = os.path.join(ROOT, "models")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def fixture_path(name):
    return os.path.join(FIXTURES, name + ".pop")


@pytest.fixture(scope="session")
def model_lib():
    """factory: compiled model library for a fixture (.pop -> CUDA -> .so, cached in odexp_b200/_build)"""
    from odexp_b200 import capi
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = capi.ModelLib.from_pop(fixture_path(name))
        return cache[name]
    return get


def load_golden(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
