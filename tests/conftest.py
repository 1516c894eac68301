import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_tests.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def nl():
    """The C-ABI library, initialised on every visible device (GPU tests only)."""
    from numpyllvm_b200 import abi
    lib = abi.load_library()
    rc = lib.nl_init(0, None)
    if rc != 0:
        pytest.fail("nl_init failed on a GPU test: " + lib.nl_last_error().decode())
    yield lib
