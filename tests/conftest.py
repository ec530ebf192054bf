import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# games whose device implementation exists (extended as games land)
IMPLEMENTED = ["maze", "coinrun", "bossfight", "climber", "chaser", "caveflyer", "jumper"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


@pytest.fixture(scope="session")
def oracle_available():
    """oracle/_ref is (re)built here when the reference sources are present; on the GPU box the
    prebuilt libraries travel with the working tree."""
    from oracle import build_ref, ref_env
    try:
        build_ref.build()
    except Exception:
        pass
    return ref_env.available()


def golden(game):
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "%s.npz" % game))
