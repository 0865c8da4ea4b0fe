import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def fixture_frame():
    """The reference's only shipped RGB-D frame (benchmark/img0.png + depth0.png), decoded."""
    import numpy as np
    z = np.load(os.path.join(ROOT, "tests", "golden", "fixture_frame.npz"))
    return z["bgr"], z["depth"]


@pytest.fixture(scope="session")
def golden():
    import json
    return json.load(open(os.path.join(ROOT, "tests", "golden", "golden_hashes.json")))
