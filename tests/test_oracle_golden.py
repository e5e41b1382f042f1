"""The CPU oracle against the reference's own known-answer vectors (src/tests/decoding.test:5-20):
sha1(decoded shown frames) must equal the vector's file name.  This is what pins the oracle."""
import hashlib
import os

import pytest

import oracle_lib as O
from conftest import GOLDEN_DIR, golden_vectors


@pytest.mark.parametrize("name", golden_vectors())
def test_oracle_reproduces_golden_sha1(name):
    data = open(os.path.join(GOLDEN_DIR, name), "rb").read()
    assert hashlib.sha1(O.decode_ivf_display(data)).hexdigest() == name


def test_all_53_vectors_present():
    assert len(golden_vectors()) == 53


def test_oracle_matches_reference_on_synthetic_1080p_and_4k_clips():
    """tests/golden/bench_clips.json holds SHA-1s of the UNMODIFIED reference's decode of the
    synthetic clips (tools/make_bench_streams.sh); the oracle must agree at these sizes too."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    want = json.load(open(os.path.join(root, "tests", "golden", "bench_clips.json")))
    for name in ("synth4k_medium_q90_8f.ivf", "synth1080p_easy_q40.ivf"):
        data = open(os.path.join(root, "bench_data", name), "rb").read()
        assert hashlib.sha1(O.decode_ivf_display(data)).hexdigest() == want[name]["sha1_of_reference_decode"], name
