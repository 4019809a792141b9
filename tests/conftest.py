import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _oracle_c_built():
    """The oracle's C restatement is test infrastructure; build it on demand."""
    so = os.path.join(ROOT, "oracle", "c", "liboracle_c.so")
    if not os.path.isfile(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle", "c")])
    yield


def golden(name):
    import json
    with open(os.path.join(ROOT, "tests", "golden", name)) as f:
        return json.load(f)
