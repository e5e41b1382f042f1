"""C++ programs written against the host mirror alfalfa_b200/host/alfalfa_gpu.hh (the reference's Decoder /
Encoder class surface over the C ABI), compiled with g++ and run on the GPU box:
  tests/cxx/decode_to_stdout.cc  the reference's src/tests/decode-to-stdout.cc shape; SHA-1 of its output
                                 must equal the vector's name (tests/decoding.test:14-15) for all 53 vectors
  tests/cxx/encoder_copies.cc    salsify-sender.cc:492-518: an Encoder copied twice per frame, both copies
                                 encoding concurrently, export_decoder / Encoder( Decoder ) round trip"""
import hashlib
import os
import subprocess

import pytest

from conftest import GOLDEN_DIR, golden_vectors

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "alfalfa_b200")


def _build(tmp_path_factory, name):
    out = str(tmp_path_factory.mktemp("cxx") / name)
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-Wall", "-Wextra", "-pthread", os.path.join(ROOT, "tests", "cxx", name + ".cc"),
                           "-o", out, "-L" + LIBDIR, "-l:libvp8gpu.so", "-Wl,-rpath," + LIBDIR])
    return out


@pytest.fixture(scope="module")
def decode_to_stdout(tmp_path_factory):
    return _build(tmp_path_factory, "decode_to_stdout")


def test_cxx_programs_compile(tmp_path_factory):
    """CPU part: both programs compile and link against the library (no device needed)"""
    _build(tmp_path_factory, "decode_to_stdout")
    _build(tmp_path_factory, "encoder_copies")


def _by_size(names):
    import oracle_lib as O
    groups = {}
    for name in names:
        w, h, _ = O.read_ivf(open(os.path.join(GOLDEN_DIR, name), "rb").read())
        groups.setdefault((w, h), []).append(name)
    return groups


@pytest.mark.gpu
@pytest.mark.parametrize("device_tokens", [0, 1])
def test_cxx_decoder_reproduces_the_golden_sha1s(decode_to_stdout, device_tokens, tmp_path):
    """one process per frame size (one Context), a fresh Decoder per vector"""
    names = [n for n in golden_vectors() if not device_tokens or n.startswith(("0", "4", "f"))]
    bad = []
    for (w, h), group in _by_size(names).items():
        out = subprocess.run([decode_to_stdout, "--out", str(tmp_path), str(device_tokens)] +
                             [os.path.join(GOLDEN_DIR, n) for n in group], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
        if out.returncode != 0:
            bad.append((w, h, out.returncode, out.stderr.decode()[-200:]))
            continue
        for n in group:
            if hashlib.sha1(open(os.path.join(str(tmp_path), n + ".yuv"), "rb").read()).hexdigest() != n:
                bad.append(n)
    assert not bad, bad


@pytest.mark.gpu
def test_cxx_encoder_copies_encode_concurrently(tmp_path_factory):
    exe = _build(tmp_path_factory, "encoder_copies")
    out = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert out.returncode == 0 and out.stdout.decode().startswith("ok"), out.stderr.decode()[-500:]
