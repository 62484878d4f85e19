import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O

    O.build()
    return O


_KEY_CACHE = {}


@pytest.fixture(scope="session")
def keyset(oracle):
    """keyset(params, seed) -> cached oracle KeySet."""

    def get(params, seed=1234, with_ksk=True):
        key = (params.name, seed, with_ksk)
        if key not in _KEY_CACHE:
            _KEY_CACHE[key] = oracle.keygen(params, seed, with_ksk=with_ksk)
        return _KEY_CACHE[key]

    return get
