import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        meta = json.load(f)
    with open(os.path.join(GOLDEN, name + ".r1cs"), "rb") as f:
        r1cs = f.read()
    with open(os.path.join(GOLDEN, name + ".pk.bin"), "rb") as f:
        pk = f.read()
    return meta, r1cs, pk


GOLDEN_NAMES = ["silly", "silly_nozk", "rand100", "rand100_circom", "rand300", "dummy924_nozk"]


@pytest.fixture(scope="session")
def gpu_ctx():
    from crescent_credentials_b200 import ffi
    ctx = ffi.Context(0)
    yield ctx
    ctx.close()
