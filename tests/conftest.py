import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "msm_vectors.json")) as f:
        return json.load(f)


def case_arrays(case):
    """(scalars [n,4], bases [n,8], result [8]) as uint64 limb arrays."""
    sc = np.frombuffer(bytes.fromhex("".join(case["scalars_mont_le"])), dtype=np.uint64).reshape(-1, 4).copy()
    bs = np.frombuffer(bytes.fromhex("".join(case["bases_mont_le"])), dtype=np.uint64).reshape(-1, 8).copy()
    res = np.frombuffer(bytes.fromhex(case["result_mont_le"]), dtype=np.uint64).copy()
    return sc, bs, res


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle

    pyoracle.build()
    return pyoracle
