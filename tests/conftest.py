import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def clip_root(tmp_path_factory):
    return str(tmp_path_factory.mktemp("clips"))


def split_instances(flat, counts, classes, keep_empty=False):
    """flat (sum,k) + per-instance counts -> list of {"class","points"} (empties dropped)."""
    out, pos = [], 0
    for cls, n in zip(classes, counts):
        n = int(n)
        if n > 0 or keep_empty:
            out.append({"class": str(cls), "points": flat[pos:pos + n]})
        pos += n
    return out
