import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sa-toolkit_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
# The grouped narrow-stage kernels are only dispatched when a launch has enough tiles to fill the GPU (small inputs take
# the per-tap kernels: less fixed cost).  The tests use small inputs, so lift the threshold: every shape a grouped plan
# exists for goes through the grouped kernels here (bench.py runs with the default threshold).
os.environ.setdefault("SATOOLS_B200_GROUP_MIN_TILES", "0")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
