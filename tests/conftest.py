import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g
    return g.load_package()


@pytest.fixture(scope="session")
def oracle(pkg):
    import __graft_entry__ as g
    o = g.load_oracle()
    o.build()
    return o


@pytest.fixture(scope="session")
def ctx(pkg):
    """CUDA context of the product library; fails loudly (no fallback) when there is no GPU."""
    return pkg._lib.default_context(0)


@pytest.fixture(scope="session")
def r1cfg(pkg):
    import numpy as np
    cfg, _ = pkg.get_default_configs()
    cfg.laser_noise = True
    cfg.tspan = np.linspace(0.0, 0.25, 11)
    return cfg
