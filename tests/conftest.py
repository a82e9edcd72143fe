import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _oracles():
    from oracle.ref import Oracle, available

    out = {}
    for kind in ("reference", "port"):
        for real in ("f64", "f32"):
            if available(kind, real):
                out[(kind, real)] = Oracle(kind, real)
    return out


@pytest.fixture(scope="session")
def oracles():
    o = _oracles()
    if not o:
        pytest.skip("no oracle library built (make -C oracle)")
    return o


@pytest.fixture(scope="session")
def oracle_f64(oracles):
    return oracles.get(("reference", "f64")) or oracles[("port", "f64")]


@pytest.fixture(scope="session")
def oracle_f32(oracles):
    return oracles.get(("reference", "f32")) or oracles[("port", "f32")]
