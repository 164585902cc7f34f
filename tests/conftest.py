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
def c1():
    """C1 fixture: the reference's bundled peptide test (164 x 54) + its golden artefacts."""
    return dict(np.load(os.path.join(GOLDEN, "peptide_c1.npz"), allow_pickle=False))


@pytest.fixture(scope="session")
def kmeans_ref():
    return dict(np.load(os.path.join(GOLDEN, "kmeans_ref.npz"), allow_pickle=False))


def synth_features(n, f, seed=0, n_slow=6, dtype=np.float32):
    """Small CPU twin of the benchmark generator: AR(1) slow modes mixed
    into F features with per-feature scale/offset (|mean| >> std)."""
    rng = np.random.default_rng(seed)
    T = 400.0 * 2.0 ** (-np.arange(n_slow))
    rho = np.exp(-1.0 / T)
    z = np.empty((n, n_slow))
    z[0] = rng.standard_normal(n_slow)
    eps = rng.standard_normal((n, n_slow))
    for t in range(1, n):
        z[t] = rho * z[t - 1] + np.sqrt(1 - rho ** 2) * eps[t]
    A = np.random.default_rng(seed + 1).standard_normal((n_slow, f))
    s = rng.uniform(0.05, 0.5, size=f)
    m = rng.uniform(0.5, 3.0, size=f)
    X = (z @ A + 0.5 * rng.standard_normal((n, f))) * s + m
    return np.ascontiguousarray(X.astype(dtype))
