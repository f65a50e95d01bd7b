import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc

    orc.build()
    return orc


@pytest.fixture(scope="session")
def ctx():
    """The GPU context; only gpu-marked tests may ask for it."""
    from raceid3_stemid2_package_b200 import lib

    c = lib.Context(device=0)
    yield c
    c.close()


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def ari(a, b):
    """Adjusted Rand index (tests only)."""
    a = np.asarray(a)
    b = np.asarray(b)
    ua, ia = np.unique(a, return_inverse=True)
    ub, ib = np.unique(b, return_inverse=True)
    tab = np.zeros((len(ua), len(ub)), dtype=np.int64)
    np.add.at(tab, (ia, ib), 1)
    comb = lambda x: x * (x - 1) / 2.0
    s = comb(tab).sum()
    sa = comb(tab.sum(1)).sum()
    sb = comb(tab.sum(0)).sum()
    n = comb(len(a))
    exp = sa * sb / n
    mx = 0.5 * (sa + sb)
    return 1.0 if mx == exp else (s - exp) / (mx - exp)
