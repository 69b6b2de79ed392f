import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as orc
    return orc


@pytest.fixture(scope="session")
def ssnamg():
    import ssnamg as s
    return s


@pytest.fixture(scope="session")
def gpu(ssnamg):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.set_device(0)
    return ssnamg


def golden_systems(prefix):
    files = sorted(f for f in os.listdir(GOLDEN) if f.startswith("ssn_system_" + prefix) and f.endswith(".npz"))
    return [os.path.join(GOLDEN, f) for f in files]


def load_system(path):
    d = dict(np.load(path))
    m, n = int(d["m"]), int(d["n"])
    d["s"] = np.unpackbits(d["s_bits"])[: m * n].astype(bool)
    d["m"], d["n"] = m, n
    return d


def random_active_problem(m, n, density, seed, weights=False):
    """A random logical active set that touches every row and column (connected-ish)."""
    rs = np.random.RandomState(seed)
    S = rs.random_sample((m, n)) < density
    S[np.arange(m), rs.randint(0, n, m)] = True
    S[rs.randint(0, m, n), np.arange(n)] = True
    p = rs.random_sample(m) + 0.5 if weights else np.ones(m)
    q = rs.random_sample(n) + 0.5 if weights else np.ones(n)
    return S.reshape(-1, order="F"), p, q
