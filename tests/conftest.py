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
    return float(np.arccos(np.clip((np.trace(np.asarray(R, dtype=np.float64)) - 1.0) / 2.0, -1.0, 1.0)))
