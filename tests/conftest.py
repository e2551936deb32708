import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import vct_b200  # noqa: E402,F401  (import shim for the hyphenated package directory)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle_py
    oracle_py.build()
    return oracle_py


@pytest.fixture()
def oracle(oracle_mod):
    o = oracle_mod.Oracle()
    yield o
    o.close()


@pytest.fixture()
def gpu_ctx():
    """A fresh context on cuda:0.  The extension is mandatory: a missing library is a failure, not a skip."""
    from vct_b200 import capi
    capi.load_library()          # raises FileNotFoundError if the .so was not built
    try:
        c = capi.Context(0)
    except capi.VctError as e:   # library present but no device: only acceptable outside `-m gpu` runs
        pytest.skip(f"no CUDA device: {e}")
    yield c
    c.close()


def psnr(a, b):
    d = a.astype(np.float64) - b.astype(np.float64)
    mse = float((d ** 2).mean())
    return 99.0 if mse == 0 else 10.0 * np.log10(255.0 ** 2 / mse)


def frac_within(a, b, tol=2):
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    if d.ndim == 3:
        d = d.max(-1)
    return float((d <= tol).mean())
