import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN, "ref_repwvl.npz"))


@pytest.fixture(scope="session")
def golden_edge():
    return np.load(os.path.join(GOLDEN, "ref_edge.npz"))


@pytest.fixture(scope="session")
def golden_misc():
    return np.load(os.path.join(GOLDEN, "ref_misc.npz"))


@pytest.fixture(scope="session")
def golden_eq():
    """16 columns x 6,000 reference iterations (tools/make_golden_equilibrium.py)."""
    return np.load(os.path.join(GOLDEN, "ref_equilibrium.npz"))


def table_path(n):
    return os.path.join(GOLDEN, f"Reduced{n}Forcing.rcmtab")


@pytest.fixture(scope="session")
def rcm():
    import our_first_climate_model_b200 as m
    if not os.path.exists(m.library_path()):
        m.build_library()
    m.load_library()
    return m


@pytest.fixture(scope="session")
def port():
    from oracle import port as P
    P.lib()
    return P
