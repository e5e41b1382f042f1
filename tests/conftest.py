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


def pytest_collection_modifyitems(config, items):
    """The re-encoding path and the two-pass (trellis) key frame were written after this round's GPU minutes were spent: its kernels and host
    code are checked bit-exactly under the SIMT emulator (tests/test_simt_emulation.py) and have had one short run on a B200
    (profiles/r2z_new_kernels_b200.json), not the whole test files.
    Its -m gpu tests therefore run LAST, so that under `pytest -x` a hardware-only failure there cannot keep the suites
    that have a hardware record (parity, encoder, C++ callers, flatten, state format) from running."""
    def is_late(it):  # (the Encoder-from-any-Decoder-state test came with the same change)
        return "reencode" in it.nodeid or "built_from_a_decoder_in_any_state" in it.nodeid or "two_pass" in it.nodeid
    late = [it for it in items if is_late(it)]
    if late:
        items[:] = [it for it in items if not is_late(it)] + late
