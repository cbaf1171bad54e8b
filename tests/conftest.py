import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_manifest():
    import json
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "manifest.json")
    with open(p) as f:
        return json.load(f)
