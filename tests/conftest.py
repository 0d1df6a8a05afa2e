import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import _pkg  # noqa: E402

_pkg.load()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O

    O.load()
    return O


@pytest.fixture(scope="session")
def hmm():
    """One library handle for the GPU tests (keeps raw FP32 sums for bit-level checks)."""
    from falcon_genome_b200 import PairHMM

    h = PairHMM(keep_raw_f32=True)
    yield h
    h.done()
