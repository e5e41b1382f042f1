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
