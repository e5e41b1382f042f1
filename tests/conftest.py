import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vectors")


def golden_vectors():
    return sorted(f for f in os.listdir(GOLDEN_DIR) if len(f) == 40)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN_DIR
