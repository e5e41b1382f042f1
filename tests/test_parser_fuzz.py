"""The host front end (alfalfa_b200/csrc/parser.cc) and the re-serialiser under AddressSanitizer + UBSan on
mutated frames of golden vectors and of the feature-complete stream: every input is answered with a status."""
import os
import subprocess

import pytest

from conftest import GOLDEN_DIR

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def fuzzer(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("fuzz") / "parser_fuzz")
    csrc = os.path.join(ROOT, "alfalfa_b200", "csrc")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                           os.path.join(ROOT, "tests", "parser_fuzz.cc"), os.path.join(csrc, "parser.cc"),
                           os.path.join(csrc, "serializer.cc"), "-pthread", "-o", exe])
    return exe


@pytest.mark.parametrize("clip,seed", [(os.path.join(GOLDEN_DIR, "0b546dad90ddefea5085c7751b5fa2f117630b1c"), 1),
                                       (os.path.join(GOLDEN_DIR, "2a4c049c2f8e3a19ee39ffd7074cecd68006a101"), 2),
                                       (os.path.join(GOLDEN_DIR, "e01c6f92f23eefecb1e120230a2c4b2767cce066"), 3),
                                       (os.path.join(ROOT, "bench_data", "features1080p_12f.ivf"), 4)])
def test_mutated_frames_are_answered_with_a_status(fuzzer, clip, seed):
    rounds = "6" if clip.endswith(".ivf") else "60"
    out = subprocess.run([fuzzer, clip, str(seed), rounds], capture_output=True, text=True, timeout=900,
                         env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0"))
    assert out.returncode == 0, out.stderr[-2000:]
    assert "parsed" in out.stdout and "rejected" in out.stdout
