import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
FAMILIES = ("nasbench101", "nasbench201", "darts")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    from nasbowl_b200.cells import CellBatch
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    batch = CellBatch.from_records(z["records"], int(z["n_max"]))
    return z, batch


def golden_kraw(z):
    """The integer Gram of a BASELINE-size fixture (stored as the packed upper triangle, u8)."""
    n = int(z["n_train"])
    K = np.zeros((n, n), dtype=np.int64)
    iu = np.triu_indices(n)
    K[iu] = z["Kraw_triu"]
    K.T[iu] = z["Kraw_triu"]
    return K


@pytest.fixture(params=FAMILIES)
def family(request):
    z, batch = load_golden(request.param)
    n = int(z["n_train"])
    return request.param, z, batch[:n], batch[n:]
