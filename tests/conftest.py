import hashlib
import json
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


def sha256(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


class Golden:
    """Fixtures generated from the compiled reference by tests/golden/make_golden.py."""

    def __init__(self):
        self.mats = np.load(os.path.join(GOLDEN, "matrices.npz"))
        self.dots = np.load(os.path.join(GOLDEN, "dots.npz"))
        with open(os.path.join(GOLDEN, "partitions.json")) as f:
            self.partitions = json.load(f)
        with open(os.path.join(GOLDEN, "systems.json")) as f:
            self.systems = json.load(f)
        self.names = sorted(self.partitions.keys())

    def csr(self, name):
        n, m, nnz = (int(v) for v in self.mats[name + ".dims"])
        return n, m, self.mats[name + ".row_ptr"], self.mats[name + ".col_ind"], self.mats[name + ".values"]

    def x(self, name):
        """x[i] = 0.25 * i, the vector of test/test_spmv.cpp:27-28."""
        return np.arange(int(self.mats[name + ".dims"][1]), dtype=np.float64) * 0.25


@pytest.fixture(scope="session")
def golden():
    return Golden()


@pytest.fixture(scope="session")
def oracle():
    from oracle import oraclebind
    oraclebind.build()
    return oraclebind


@pytest.fixture(scope="session")
def gpu_lib():
    """The C-ABI library on a machine with a GPU. Fails (does not skip) if it cannot compute."""
    import cask_b200
    cask_b200.lib()
    return cask_b200


@pytest.fixture()
def ctx(gpu_lib):
    c = gpu_lib.Context(0)
    yield c
    c.close()


def row_scale(n, row_ptr, col_ind, values, x):
    """sum_j |a_ij x_j| per row: the magnitude an fp64 row sum is accurate relative to."""
    rows = np.repeat(np.arange(n), np.diff(row_ptr))
    return np.bincount(rows, weights=np.abs(values * x[col_ind]), minlength=n)


def assert_y_close(got, exp, scale, tol=1e-12):
    """north_star tolerance: 1e-12 relative per entry. The relative error is taken against
    max(|exp_i|, sum_j |a_ij x_j|): equal to |exp_i| for rows without cancellation, and the only
    meaningful yardstick for rows whose terms cancel (a different summation order moves the
    result by ~1e-16 * sum|terms|, whatever |exp_i| happens to be)."""
    den = np.maximum(np.abs(exp), scale)
    err = np.abs(got - exp)
    bad = err > tol * den
    assert not bad.any(), "max rel err %.3e at %d (got %r exp %r)" % (
        (err / np.maximum(den, 1e-300)).max(), int(np.argmax(err / np.maximum(den, 1e-300))),
        got[np.argmax(bad)], exp[np.argmax(bad)])
    # the reference's own check, test/test_utils.hpp:36: almost_equal(got, exp, 1E-8, 1E-11)
    assert np.all(err <= 1e-8 * np.abs(exp) + 1e-11 * np.maximum(1.0, scale))
