import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def fixtures():
    return dict(np.load(os.path.join(GOLDEN, "fixtures.npz")))


@pytest.fixture(scope="session")
def cv2_golden():
    import json
    g = dict(np.load(os.path.join(GOLDEN, "cv2_golden.npz")))
    g["meta"] = json.loads(str(g["meta_json"]))
    return g


@pytest.fixture(scope="session")
def calib(fixtures):
    return {s: {k: fixtures["%s_%s" % (s, k)] for k in "KDRP"} for s in ("left", "right")}
