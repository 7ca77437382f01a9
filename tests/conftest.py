import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


_PAIR_CACHE = {}


@pytest.fixture(scope="session")
def pair_cache():
    """make_pair results memoised per session (C3 ray casting takes ~15 s at full size)."""
    from aicp_mapping_b200 import synth

    def get(config, trial=0, n_points=None):
        key = (config, trial, n_points)
        if key not in _PAIR_CACHE:
            _PAIR_CACHE[key] = synth.make_pair(config, trial, n_points)
        return _PAIR_CACHE[key]
    return get


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.lib()
    return oracle


def rot_angle(R):
    """Rotation angle of a (nearly) orthonormal 3x3, well conditioned near zero: atan2(|skew part|, (trace - 1) / 2)."""
    R = np.asarray(R, dtype=np.float64)
    v = 0.5 * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    return float(np.arctan2(np.linalg.norm(v), (np.trace(R) - 1.0) / 2.0))
