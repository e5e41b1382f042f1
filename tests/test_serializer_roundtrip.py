"""Byte-exact re-serialisation gate (the reference's src/tests/roundtrip.cc:93-112 / tests/roundtrip-verify.test):
every frame of every golden vector is parsed by the product's front end (vp8gpu_parse_frame with the labels
kept) and written back by its bitstream writer (vp8gpu_parsed_serialize = Frame::serialize,
encoder/serializer.cc:388-405); the output must equal the input frame byte for byte.  This pins the writer's
bool coder, header syntax, mode / vector coding with census contexts, SPLITMV labels, token coding with
its contexts and the partition layout (1-8 DCT partitions, segmentation, golden / altref, probability
updates) against streams the product did not write (libvpx's, via the reference's vectors).
Also run over the feature-complete synthetic streams and the reference encoder's bench clips."""
import ctypes as C
import os
import sys

import pytest

import oracle_lib as O
from alfalfa_b200 import capi
from conftest import GOLDEN_DIR, golden_vectors

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _roundtrip(data, max_frames=None):
    L = capi.lib()
    w, h, frames = O.read_ivf(data)
    st, pf = C.c_void_p(), C.c_void_p()
    capi.check(L.vp8gpu_state_create(w, h, C.byref(st)))
    capi.check(L.vp8gpu_parsed_create(C.byref(pf)))
    capi.check(L.vp8gpu_parsed_keep_labels(pf, 1))
    out = (C.c_uint8 * (max(len(f) for f in frames) + 4096))()
    try:
        for i, f in enumerate(frames[:max_frames]):
            assert L.vp8gpu_parse_frame(st, f, len(f), pf) == 0, "frame %d does not parse" % i
            size = C.c_size_t(0)
            assert L.vp8gpu_parsed_serialize(pf, out, len(out), C.byref(size)) == 0, "frame %d" % i
            assert size.value == len(f), "frame %d: %d bytes written, %d read" % (i, size.value, len(f))
            assert C.string_at(out, size.value) == f, "frame %d differs" % i
    finally:
        L.vp8gpu_state_destroy(st)
        L.vp8gpu_parsed_destroy(pf)
    return len(frames[:max_frames])


@pytest.mark.parametrize("name", golden_vectors())
def test_golden_vector_reserialises_byte_for_byte(name):
    assert _roundtrip(open(os.path.join(GOLDEN_DIR, name), "rb").read()) >= 1


@pytest.mark.parametrize("clip", ["features1080p_12f.ivf", "synth1080p_medium_q90.ivf", "synth720p_medium_q90.ivf",
                                  "synth4k_medium_q90_8f.ivf"])
def test_bench_clip_reserialises_byte_for_byte(clip):
    """feature-complete stream (own writer, all features) and streams written by the reference encoder"""
    assert _roundtrip(open(os.path.join(ROOT, "bench_data", clip), "rb").read(), 12) >= 8


@pytest.mark.parametrize("w,h,seed", [(176, 144, 11), (175, 143, 12), (33, 17, 13), (640, 368, 14)])
def test_fresh_feature_stream_reserialises_byte_for_byte(w, h, seed):
    import make_feature_stream
    assert _roundtrip(make_feature_stream.make_stream(w, h, 8, seed)) == 8


def test_serialize_needs_kept_labels():
    L = capi.lib()
    name = golden_vectors()[0]
    w, h, frames = O.read_ivf(open(os.path.join(GOLDEN_DIR, name), "rb").read())
    st, pf = C.c_void_p(), C.c_void_p()
    capi.check(L.vp8gpu_state_create(w, h, C.byref(st)))
    capi.check(L.vp8gpu_parsed_create(C.byref(pf)))
    assert L.vp8gpu_parse_frame(st, frames[0], len(frames[0]), pf) == 0
    size = C.c_size_t(0)
    assert L.vp8gpu_parsed_serialize(pf, None, 0, C.byref(size)) == capi.ERR_LOGIC
    capi.check(L.vp8gpu_parsed_keep_labels(pf, 1))
    st2 = C.c_void_p()
    capi.check(L.vp8gpu_state_create(w, h, C.byref(st2)))
    assert L.vp8gpu_parse_frame(st2, frames[0], len(frames[0]), pf) == 0
    assert L.vp8gpu_parsed_serialize(pf, None, 0, C.byref(size)) == capi.ERR_NOMEM and size.value == len(frames[0])
    for s in (st, st2):
        L.vp8gpu_state_destroy(s)
    L.vp8gpu_parsed_destroy(pf)
