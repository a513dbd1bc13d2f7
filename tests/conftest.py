import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def graph_from_npz(d):
    """Rebuild the reference's N x N match lists from a golden file (see tests/golden/make_golden.py)."""
    n = int(d["n_kf"])
    points = [d["points_%d" % i] for i in range(n)]
    src = [[[] for _ in range(n)] for _ in range(n)]
    dst = [[[] for _ in range(n)] for _ in range(n)]
    lmk = [[[] for _ in range(n)] for _ in range(n)]
    for i, j in d["pairs"]:
        src[i][j] = d["src_%d_%d" % (i, j)].tolist()
        dst[i][j] = d["dst_%d_%d" % (i, j)].tolist()
        lmk[i][j] = d["lmk_%d_%d" % (i, j)].tolist()
    return points, src, dst, lmk, int(d["n_landmark"])


@pytest.fixture(scope="session")
def golden():
    return load_golden
