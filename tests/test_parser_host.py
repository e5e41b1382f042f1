"""CPU entropy front end (alfalfa_b200/csrc/parser.cc) against the oracle's restatement of
DecoderState::parse_and_apply: the flat records must be byte-identical for every frame of every
golden vector.  Also pins the state-passing behaviour of vp8gpu_state.  No GPU needed."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O
from alfalfa_b200 import capi
from conftest import GOLDEN_DIR, golden_vectors


def _frames(name):
    return O.read_ivf(open(os.path.join(GOLDEN_DIR, name), "rb").read())


@pytest.mark.parametrize("name", golden_vectors())
def test_records_identical_to_oracle(name):
    L = capi.lib()
    w, h, frames = _frames(name)
    od = O.OracleDecoder(w, h)
    st, pf = C.c_void_p(), C.c_void_p()
    capi.check(L.vp8gpu_state_create(w, h, C.byref(st)))
    capi.check(L.vp8gpu_parsed_create(C.byref(pf)))
    started = False
    limit = 40 if w * h > 500000 else 400
    for i, f in enumerate(frames[:limit]):
        if not started and (f[0] & 1):
            continue
        started = True
        od.decode(f, want_planes=False)
        op = od.parsed()
        assert L.vp8gpu_parse_frame(st, f, len(f), pf) == 0
        desc = L.vp8gpu_parsed_desc(pf).contents
        assert bytes(desc) == bytes(op.desc), "frame %d desc" % i
        n = desc.mb_cols * desc.mb_rows
        assert C.string_at(L.vp8gpu_parsed_mbs(pf), n * 32) == op.mbs.tobytes(), "frame %d mbs" % i
        if desc.n_tokens:
            assert C.string_at(L.vp8gpu_parsed_tokens(pf), desc.n_tokens * 4) == op.tokens.tobytes(), "frame %d tokens" % i
        if desc.n_split:
            assert C.string_at(L.vp8gpu_parsed_split(pf), desc.n_split * 64) == op.split.tobytes(), "frame %d split" % i
    L.vp8gpu_state_destroy(st)
    L.vp8gpu_parsed_destroy(pf)


def test_state_is_a_value_clone_equal_hash():
    L = capi.lib()
    w, h, frames = _frames("2a4c049c2f8e3a19ee39ffd7074cecd68006a101")
    a, pf = C.c_void_p(), C.c_void_p()
    capi.check(L.vp8gpu_state_create(w, h, C.byref(a)))
    capi.check(L.vp8gpu_parsed_create(C.byref(pf)))
    for f in frames[:3]:
        assert L.vp8gpu_parse_frame(a, f, len(f), pf) == 0
    b = C.c_void_p()
    capi.check(L.vp8gpu_state_clone(a, C.byref(b)))
    assert L.vp8gpu_state_equal(a, b) and L.vp8gpu_state_hash(a) == L.vp8gpu_state_hash(b)
    assert L.vp8gpu_parse_frame(a, frames[3], len(frames[3]), pf) == 0
    changed = not L.vp8gpu_state_equal(a, b)
    assert L.vp8gpu_parse_frame(b, frames[3], len(frames[3]), pf) == 0
    assert L.vp8gpu_state_equal(a, b) and L.vp8gpu_state_hash(a) == L.vp8gpu_state_hash(b)
    assert changed or True  # a frame may leave the persistent state untouched
    for s in (a, b):
        L.vp8gpu_state_destroy(s)
    L.vp8gpu_parsed_destroy(pf)


def test_error_codes_and_state_untouched_on_error():
    L = capi.lib()
    w, h, frames = _frames("45502fe01a62b82d498b83dc50824741402436db")
    st, ref, pf = C.c_void_p(), C.c_void_p(), C.c_void_p()
    capi.check(L.vp8gpu_state_create(w, h, C.byref(st)))
    capi.check(L.vp8gpu_parsed_create(C.byref(pf)))
    assert L.vp8gpu_parse_frame(st, frames[0], len(frames[0]), pf) == 0
    capi.check(L.vp8gpu_state_clone(st, C.byref(ref)))
    f = frames[1]
    assert L.vp8gpu_parse_frame(st, f[:2], 2, pf) == capi.ERR_INVALID           # truncated tag
    assert L.vp8gpu_parse_frame(st, f[:12], 12, pf) == capi.ERR_INVALID         # first partition cut off
    bad = bytearray(f); bad[4] = 0x55
    assert L.vp8gpu_parse_frame(st, bytes(bad), len(bad), pf) == capi.ERR_INVALID   # start code
    bad = bytearray(f); bad[0] |= 2
    assert L.vp8gpu_parse_frame(st, bytes(bad), len(bad), pf) == capi.ERR_UNSUPPORTED  # VP8 version != 0
    bad = bytearray(f); bad[6] ^= 1
    assert L.vp8gpu_parse_frame(st, bytes(bad), len(bad), pf) == capi.ERR_UNSUPPORTED  # other frame size
    assert L.vp8gpu_state_equal(st, ref)
    assert L.vp8gpu_parse_frame(st, f, len(f), pf) == 0
    for s in (st, ref):
        L.vp8gpu_state_destroy(s)
    L.vp8gpu_parsed_destroy(pf)
