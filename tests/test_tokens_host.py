"""Device-side token decoder (alfalfa_b200/csrc/tokens_core.cuh, the body of k_tokens) compiled for
the CPU: together with parse_frame(defer_tokens) it must reproduce, byte for byte, the records of
the CPU front end (parser.cc, itself pinned to the oracle by test_parser_host.py) on every frame
of every golden vector -- Frame::parse_tokens (frame.cc:122-137), Block::parse_tokens
(tokens.cc:50-135), BoolDecoder (bool_decoder.hh:82-107).  Also checks the token-pool capacity
rule of Engine::token_ring_create and the behaviour on truncated partitions.  No GPU needed."""
import ctypes as C
import os
import subprocess
import tempfile

import pytest

import oracle_lib as O
from conftest import GOLDEN_DIR, golden_vectors

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "alfalfa_b200", "csrc")


@pytest.fixture(scope="module")
def shim():
    d = tempfile.mkdtemp()
    so = os.path.join(d, "tokens_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall",
                           os.path.join(ROOT, "tests", "tokens_host_shim.cc"), os.path.join(CSRC, "parser.cc"),
                           "-o", so])
    L = C.CDLL(so)
    L.th_new.restype = C.c_void_p
    L.th_new.argtypes = [C.c_int, C.c_int]
    L.th_free.argtypes = [C.c_void_p]
    L.th_variant.argtypes = [C.c_void_p, C.c_int]
    L.th_frame.restype = C.c_int
    L.th_frame.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_uint32, C.POINTER(C.c_uint32)]
    return L


@pytest.mark.parametrize("lockstep", [0, 1])
@pytest.mark.parametrize("name", golden_vectors())
def test_device_token_logic_matches_cpu_front_end(shim, name, lockstep):
    """both forms of the kernel body: one thread per frame, and the one-decision-per-iteration state machine"""
    w, h, frames = O.read_ivf(open(os.path.join(GOLDEN_DIR, name), "rb").read())
    H = shim.th_new(w, h)
    shim.th_variant(H, lockstep)
    started = False
    limit = 40 if w * h > 500000 else 400
    n = C.c_uint32(0)
    for i, f in enumerate(frames[:limit]):
        if not started and (f[0] & 1):
            continue
        started = True
        rc = shim.th_frame(H, f, len(f), 0, C.byref(n))
        assert rc == 0, "frame %d: code %d" % (i, rc)
        # pool rule: at most ~8.1 tokens per partition byte
        assert n.value <= 9 * len(f) + 1024
    shim.th_free(H)


def test_truncated_partitions(shim):
    """Past the end of a partition the decoder sees zero bits (bool_decoder.hh:95-99): both front
    ends must agree on what that yields, for cuts everywhere in the token data."""
    name = golden_vectors()[0]
    w, h, frames = O.read_ivf(open(os.path.join(GOLDEN_DIR, name), "rb").read())
    f = next(fr for fr in frames if not (fr[0] & 1))
    first = ((f[0] | (f[1] << 8) | (f[2] << 16)) >> 5) + 10
    for cut in list(range(first + 1, min(len(f), first + 40))) + [len(f) - 1, len(f) - 7, (first + len(f)) // 2]:
        H = shim.th_new(w, h)
        shim.th_variant(H, cut & 1)
        rc = shim.th_frame(H, f[:cut], cut, 0, None)
        assert rc <= 0, "cut %d: code %d" % (cut, rc)
        shim.th_free(H)
