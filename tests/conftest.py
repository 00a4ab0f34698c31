import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py
    oracle_py.build()
    return oracle_py


def golden_path(name):
    return os.path.join(ROOT, "tests", "golden", name)


def assert_history_1e12(hist, ref, k, label=""):
    """north_star tolerance for fp64 residuals: every entry of the first k (<= one restart cycle) within 1e-12 of the
    reference RELATIVE TO THE INITIAL RESIDUAL.  Prints the measured margin (pytest -s / -rA shows it)."""
    import numpy as np
    h, r = np.asarray(hist[:k], dtype=float), np.asarray(ref[:k], dtype=float)
    dev = float(np.max(np.abs(h - r))) / float(r[0]) if k else 0.0
    print("history %s: max|h-ref|/r0 = %.2e over %d entries (bound 1e-12, margin %.0fx)" % (label, dev, k, 1e-12 / dev if dev else float("inf")))
    assert dev <= 1e-12, (label, dev)


SF_OPS = ("replace", "sum", "prod", "max", "min")


def sf_graph_order(g):
    """The order the reference applies the leaves of a PetscSF fixture in: PetscSFSetGraph sorts them by location
    (sf.c:500; locations are distinct) and PETSCSFBASIC keeps that order for the process-local part."""
    import numpy as np
    perm = np.argsort(g["local"], kind="stable")
    return g["local"][perm].astype(np.int32), g["remote"][perm].astype(np.int32)
