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
    """The CPU oracle (test infrastructure): oracle/oracle.c through ctypes."""
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def ctx():
    """A b200zkp context on cuda:0; only gpu-marked tests may request it."""
    import intmax_zkp_core_b200 as z
    c = z.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def golden():
    import json
    d = os.path.join(ROOT, "tests", "golden")
    out = {}
    for name in ("reference_poseidon_kats", "commit_vectors", "opening_proof_vectors"):
        with open(os.path.join(d, name + ".json")) as f:
            out[name] = json.load(f)
    return out
