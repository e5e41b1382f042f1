"""Bitstream writer (alfalfa_b200/csrc/serializer.cc) on the CPU.
Round trip through two independent readers: records parsed from the golden vectors are
re-serialised (our own header: no segmentation, default probabilities), then parsed again by
(a) the product parser and (b) the oracle's restatement of the reference parser; modes, motion
vectors, B_PRED modes and every quantised coefficient must survive.  The reference has the same
kind of check (src/tests/roundtrip.cc:93-112), there byte-exact because it re-uses the original
header."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O
from alfalfa_b200 import capi
from conftest import GOLDEN_DIR, golden_vectors


def _roundtrip_vector(name, max_frames, optimize):
    L = capi.lib()
    w, h, frames = O.read_ivf(open(os.path.join(GOLDEN_DIR, name), "rb").read())
    st, pf = C.c_void_p(), C.c_void_p()
    capi.check(L.vp8gpu_state_create(w, h, C.byref(st)))
    capi.check(L.vp8gpu_parsed_create(C.byref(pf)))
    st2, pf2 = C.c_void_p(), C.c_void_p()  # state fed only with re-serialised frames
    capi.check(L.vp8gpu_state_create(w, h, C.byref(st2)))
    capi.check(L.vp8gpu_parsed_create(C.byref(pf2)))
    od = O.OracleDecoder(w, h)           # the oracle also decodes the re-serialised stream
    out = (C.c_uint8 * (4 << 20))()
    done = skipped = 0
    started = False
    for f in frames[:max_frames]:
        if not started and (f[0] & 1):
            continue
        started = True
        assert L.vp8gpu_parse_frame(st, f, len(f), pf) == 0
        d = L.vp8gpu_parsed_desc(pf).contents
        n = d.mb_cols * d.mb_rows
        mbs = np.frombuffer(C.string_at(L.vp8gpu_parsed_mbs(pf), n * 32), dtype=capi.MB_DTYPE).copy()
        hdr = capi.EncodeHeader(width=w, height=h, key_frame=d.key_frame, show_frame=d.show_frame, y_ac_qi=40,
                                loop_filter_level=d.loop_filter_level, sharpness=d.sharpness,
                                optimize_token_probs=int(optimize))
        size = C.c_size_t(0)
        rc = L.vp8gpu_serialize_frame(C.byref(hdr), L.vp8gpu_parsed_mbs(pf), L.vp8gpu_parsed_tokens(pf),
                                      L.vp8gpu_parsed_split(pf), out, len(out), C.byref(size))
        if rc == capi.ERR_UNSUPPORTED:
            # golden / altref references are outside the written subset: the chain of re-serialised
            # frames ends here (later frames would predict from a different reference set)
            skipped += 1
            break
        assert rc == 0
        blob = bytes(out[:size.value])
        assert L.vp8gpu_parse_frame(st2, blob, len(blob), pf2) == 0
        d2 = L.vp8gpu_parsed_desc(pf2).contents
        mbs2 = np.frombuffer(C.string_at(L.vp8gpu_parsed_mbs(pf2), n * 32), dtype=capi.MB_DTYPE)
        od.decode(blob, want_planes=False)
        op = od.parsed()
        for got in (mbs2, op.mbs):
            assert np.array_equal(got["ref_frame"], mbs["ref_frame"])
            assert np.array_equal(got["tok_cnt"], mbs["tok_cnt"]) and np.array_equal(got["tok_off"], mbs["tok_off"])
            assert np.array_equal(got["uv_mode"], mbs["uv_mode"]) and np.array_equal(got["b_modes"], mbs["b_modes"])
            assert np.array_equal(got["mv_x"], mbs["mv_x"]) and np.array_equal(got["mv_y"], mbs["mv_y"])
            intra = mbs["ref_frame"] == 0
            assert np.array_equal(got["y_mode"][intra], mbs["y_mode"][intra])
            assert np.array_equal(got["y_mode"] == 9, mbs["y_mode"] == 9)  # SPLITMV stays SPLITMV
        assert d2.n_tokens == d.n_tokens == op.desc.n_tokens and d2.n_split == d.n_split
        if d.n_tokens:
            t = C.string_at(L.vp8gpu_parsed_tokens(pf), d.n_tokens * 4)
            assert C.string_at(L.vp8gpu_parsed_tokens(pf2), d.n_tokens * 4) == t == op.tokens.tobytes()
        if d.n_split:
            s = C.string_at(L.vp8gpu_parsed_split(pf), d.n_split * 64)
            assert C.string_at(L.vp8gpu_parsed_split(pf2), d.n_split * 64) == s == op.split.tobytes()
        done += 1
    for x in (st, st2):
        L.vp8gpu_state_destroy(x)
    for x in (pf, pf2):
        L.vp8gpu_parsed_destroy(x)
    return done, skipped


@pytest.mark.parametrize("name", golden_vectors())
def test_reserialised_frames_parse_back_identically(name):
    done, skipped = _roundtrip_vector(name, 30 if name.startswith("ff29") else 120, optimize=False)
    assert done >= 1


@pytest.mark.parametrize("name", ["2a4c049c2f8e3a19ee39ffd7074cecd68006a101", "e01c6f92f23eefecb1e120230a2c4b2767cce066",
                                  "45502fe01a62b82d498b83dc50824741402436db"])
def test_roundtrip_with_optimised_token_probabilities(name):
    done, _ = _roundtrip_vector(name, 40, optimize=True)
    assert done >= 1


def test_writer_threads_under_thread_sanitizer(tmp_path):
    """hostpool.h + the writer's parallel sections + the late loop-filter level (tests/serializer_threads.cc) under
    -fsanitize=thread: no data race, same bytes as single-threaded"""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    clip = os.path.join(root, "bench_data", "synth1080p_medium_q90.ivf")
    if shutil.which("g++") is None or not os.path.exists(clip):
        pytest.skip("needs g++ and bench_data")
    src = os.path.join(root, "alfalfa_b200", "csrc")
    exe = str(tmp_path / "serializer_threads")
    r = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=thread", "-pthread", "-I", src, "-I", os.path.join(root, "include"),
                        os.path.join(root, "tests", "serializer_threads.cc"), os.path.join(src, "parser.cc"), os.path.join(src, "serializer.cc"),
                        "-o", exe], capture_output=True, text=True, timeout=600)
    if r.returncode != 0 and "tsan" in (r.stderr or "").lower():
        pytest.skip("ThreadSanitizer runtime not available")
    assert r.returncode == 0, r.stderr[-1500:]
    r = subprocess.run([exe, clip], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "bad 0" in r.stdout and "ThreadSanitizer" not in r.stderr, (r.stdout + r.stderr)[-2000:]
