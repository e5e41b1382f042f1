"""GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI, against
(a) the reference's golden SHA-1 vectors and (b) the CPU oracle frame by frame.  Bit-exact:
this is integer / byte work, the tolerance is zero."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

import oracle_lib as O
from conftest import GOLDEN_DIR, golden_vectors

pytestmark = pytest.mark.gpu


def _read(name):
    return open(os.path.join(GOLDEN_DIR, name), "rb").read()


@pytest.mark.parametrize("name", golden_vectors())
def test_fileplayer_reproduces_golden_sha1(name):
    """decode-to-stdout (src/tests/decode-to-stdout.cc:43-49) on the GPU path: sha1 == file name"""
    from alfalfa_b200 import Context, FilePlayer
    data = _read(name)
    w, h, _ = O.read_ivf(data)
    ctx = Context(w, h, max_frames=16)
    player = FilePlayer(ctx, data)
    sha = hashlib.sha1()
    while not player.eof():
        sha.update(player.advance().display_bytes())
    del player
    ctx.close()
    assert sha.hexdigest() == name


@pytest.mark.parametrize("device_tokens", [False, True])
@pytest.mark.parametrize("threads", [1, 4])
@pytest.mark.parametrize("name", ["2a4c049c2f8e3a19ee39ffd7074cecd68006a101",
                                  "45502fe01a62b82d498b83dc50824741402436db",
                                  "ff2941dde20090835032c32c0644b6d401610c57"])
def test_gop_parallel_stream_decode_matches_golden(name, threads, device_tokens):
    """vp8gpu_decode_ivf, DCT partitions decoded by the host workers or by k_tokens on the device"""
    from alfalfa_b200 import Context, decode_ivf
    data = _read(name)
    w, h, _ = O.read_ivf(data)
    ctx = Context(w, h, max_frames=48)
    ctx.set_device_tokens(device_tokens)
    out, n_dec, n_shown = decode_ivf(ctx, data, threads=threads)
    ctx.close()
    assert n_shown > 0 and hashlib.sha1(out).hexdigest() == name


@pytest.mark.parametrize("name", ["0b546dad90ddefea5085c7751b5fa2f117630b1c",
                                  "e01c6f92f23eefecb1e120230a2c4b2767cce066",
                                  "a4dace04a77fc9f969a8d7a645c99c0271f1f73e"])
def test_every_frame_matches_oracle_including_hidden_and_references(name):
    """frame-by-frame (not only shown frames): decoded raster and the three references"""
    from alfalfa_b200 import Context, Decoder
    data = _read(name)
    w, h, frames = O.read_ivf(data)
    ctx = Context(w, h, max_frames=16)
    dec = Decoder(ctx)
    od = O.OracleDecoder(w, h)
    started = False
    for i, f in enumerate(frames[:40]):
        if not started and (f[0] & 1):
            continue
        started = True
        want = od.decode(f)
        shown, raster = dec.get_frame_output(f)
        assert shown == want["shown"]
        for g, w_ in zip(raster.planes(), want["planes"]):
            assert np.array_equal(g, w_), "frame %d" % i
        raster.release()
    for k, r in enumerate(dec.get_references()):
        for g, w_ in zip(r.planes(), O.raster_planes(od.L.vp8o_decoder_ref(od.d, k))):
            assert np.array_equal(g, w_), "reference %d" % k
        r.release()
    del dec
    ctx.close()


def test_decoder_copy_shares_state_and_diverges_independently():
    """explicit state passing: a copied Decoder continues independently (salsify-sender.cc:492-518)"""
    from alfalfa_b200 import Context, Decoder
    data = _read("2a4c049c2f8e3a19ee39ffd7074cecd68006a101")
    w, h, frames = O.read_ivf(data)
    ctx = Context(w, h, max_frames=24)
    a = Decoder(ctx)
    for f in frames[:5]:
        a.get_frame_output(f)[1].release()
    b = a.copy()
    assert a == b
    ra = a.get_frame_output(frames[5])[1]
    assert not (a == b)
    rb = b.get_frame_output(frames[5])[1]
    assert a == b
    assert all(np.array_equal(x, y) for x, y in zip(ra.planes(), rb.planes()))
    ra.release()
    rb.release()
    del a, b
    ctx.close()


def test_batched_seam_matches_oracle():
    """vp8gpu_decode_batch: independent key frames of one stream decoded in one set of launches"""
    from alfalfa_b200 import Context, capi
    data = _read("45502fe01a62b82d498b83dc50824741402436db")  # 320x240, 30 key frames
    w, h, frames = O.read_ivf(data)
    ctx = Context(w, h, max_frames=40)
    L = ctx.L
    od = O.OracleDecoder(w, h)
    keep, jobs, outs, wants = [], (capi.Job * 8)(), [], []
    for i, f in enumerate(frames[:8]):
        r = od.decode(f)
        p = od.parsed()
        assert p.desc.key_frame
        mbs, tok, sp = (np.ascontiguousarray(x) for x in (p.mbs, p.tokens, p.split))
        desc = capi.FrameDesc.from_buffer_copy(bytes(p.desc))
        out = ctx.alloc_frame()
        keep.append((mbs, tok, sp, desc))
        jobs[i].desc = C.pointer(desc)
        jobs[i].mbs = mbs.ctypes.data
        jobs[i].tokens = tok.ctypes.data if tok.size else None
        jobs[i].split = None
        jobs[i].refs[:] = [-1, -1, -1]
        jobs[i].out = out.id
        outs.append(out)
        wants.append(r["planes"])
    capi.check(L.vp8gpu_decode_batch(ctx.h, 0, jobs, 8), ctx.h, "decode_batch")
    for out, want in zip(outs, wants):
        assert all(np.array_equal(g, w_) for g, w_ in zip(out.planes(), want))
        out.release()
    ctx.close()


def test_errors_mirror_reference_exception_classes():
    from alfalfa_b200 import Context, Decoder, Invalid, Unsupported
    ctx = Context(320, 240)
    dec = Decoder(ctx)
    with pytest.raises(Invalid):
        dec.get_frame_output(b"\x00\x00")  # truncated tag
    data = _read("45502fe01a62b82d498b83dc50824741402436db")
    _, _, frames = O.read_ivf(data)
    bad = bytearray(frames[0])
    bad[3] = 0  # broken start code (uncompressed_chunk.cc:101-103)
    with pytest.raises(Invalid):
        dec.get_frame_output(bytes(bad))
    other = Context(176, 144)
    with pytest.raises(Unsupported):  # size mismatch (uncompressed_chunk.cc:111-115)
        Decoder(other).get_frame_output(frames[0])
    del dec
    ctx.close()
    other.close()


@pytest.mark.parametrize("device_tokens", [False, True])
@pytest.mark.parametrize("name,threads", [("synth1080p_medium_q90.ivf", 8), ("synth1080p_easy_q40.ivf", 3),
                                          ("synth4k_medium_q90_8f.ivf", 2), ("features1080p_12f.ivf", 1),
                                          ("synth720p_medium_q90.ivf", 2)])
def test_full_size_clips_match_reference_decode(name, threads, device_tokens):
    """BASELINE.json sizes (1080p bench workload, 4K, and the feature-complete 1080p stream of
    tools/make_feature_stream.py): GPU decode through vp8gpu_decode_ivf vs the SHA-1 of the unmodified
    reference's decode of the same clip (tests/golden/bench_clips.json)."""
    import json
    from alfalfa_b200 import Context, decode_ivf
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    want = json.load(open(os.path.join(root, "tests", "golden", "bench_clips.json")))[name]
    data = open(os.path.join(root, "bench_data", name), "rb").read()
    w, h, _ = O.read_ivf(data)
    ctx = Context(w, h, max_frames=256 if device_tokens else 64)
    ctx.set_device_tokens(device_tokens)
    out, n_dec, n_shown = decode_ivf(ctx, data, threads=threads)
    ctx.close()
    assert len(out) == want["bytes"] and hashlib.sha1(out).hexdigest() == want["sha1_of_reference_decode"]


def test_many_independent_720p_streams_in_one_batch():
    """BASELINE.json config 5 shape: frames of independent streams decoded by one batched launch set
    (here 12 key frames of the real 720p vector as 12 streams)."""
    from alfalfa_b200 import Context, capi
    data = _read("ff2941dde20090835032c32c0644b6d401610c57")
    w, h, frames = O.read_ivf(data)
    ctx = Context(w, h, max_frames=32)
    L = ctx.L
    n = 12
    keep, jobs, outs, wants = [], (capi.Job * n)(), [], []
    for i in range(n):
        od = O.OracleDecoder(w, h)  # every stream has its own decoder state
        r = od.decode(frames[i * 5])
        p = od.parsed()
        mbs, tok = np.ascontiguousarray(p.mbs), np.ascontiguousarray(p.tokens)
        desc = capi.FrameDesc.from_buffer_copy(bytes(p.desc))
        out = ctx.alloc_frame()
        keep.append((mbs, tok, desc))
        jobs[i].desc, jobs[i].mbs, jobs[i].tokens, jobs[i].split = C.pointer(desc), mbs.ctypes.data, tok.ctypes.data, None
        jobs[i].refs[:] = [-1, -1, -1]
        jobs[i].out = out.id
        outs.append(out)
        wants.append(r["planes"])
    capi.check(L.vp8gpu_decode_batch(ctx.h, 0, jobs, n), ctx.h, "decode_batch")
    for out, want in zip(outs, wants):
        assert all(np.array_equal(g, w_) for g, w_ in zip(out.planes(), want))
        out.release()
    ctx.close()


def test_decoder_hash_follows_equality():
    """Decoder::get_hash / minihash: copies and independently built equal decoders hash equally, a diverged
    decoder does not (decoder.hh:279-292)"""
    from alfalfa_b200 import Context, Decoder
    data = _read("0b546dad90ddefea5085c7751b5fa2f117630b1c")
    w, h, frames = O.read_ivf(data)
    ctx = Context(w, h, max_frames=32)
    a, b = Decoder(ctx), Decoder(ctx)
    assert a.get_hash() == b.get_hash()
    for f in frames[:6]:
        a.get_frame_output(f)[1].release()
    assert a.get_hash() != b.get_hash()
    for f in frames[:6]:
        b.get_frame_output(f)[1].release()
    assert a == b and a.get_hash() == b.get_hash() and a.minihash() == b.minihash()
    c = a.copy()
    assert c.get_hash() == a.get_hash()
    c.get_frame_output(frames[6])[1].release()
    assert c.get_hash() != a.get_hash()
    del a, b, c
    ctx.close()


def test_device_hash_tracks_content_and_references():
    """RasterHandle::hash analogue: equal rasters hash equal, different ones differ; the three
    references of two decoders fed the same frames hash identically (multi-GPU correctness check)."""
    from alfalfa_b200 import Context, Decoder
    data = _read("2a4c049c2f8e3a19ee39ffd7074cecd68006a101")
    w, h, frames = O.read_ivf(data)
    ctx = Context(w, h, max_frames=32)
    a, b = Decoder(ctx), Decoder(ctx)
    hashes = []
    for f in frames[:6]:
        ra, rb = a.get_frame_output(f)[1], b.get_frame_output(f)[1]
        assert ra.hash() == rb.hash()
        hashes.append(ra.hash())
        ra.release()
        rb.release()
    assert len(set(hashes)) == len(hashes)
    for x, y in zip(a.get_references(), b.get_references()):
        assert x.hash() == y.hash()
        x.release()
        y.release()
    del a, b
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("device_tokens", [False, True])
def test_decode_ivf_unwinds_after_a_corrupt_frame_in_the_middle(device_tokens):
    """A frame that does not parse in GOP k of many: vp8gpu_decode_ivf returns the parser's error, every worker and
    dispatcher retires, every raster goes back to the pool, and the same context then decodes the intact stream
    bit-exactly."""
    from alfalfa_b200 import Context, capi, decode_ivf, write_ivf
    name = "0b546dad90ddefea5085c7751b5fa2f117630b1c"
    data = open(os.path.join(GOLDEN_DIR, name), "rb").read()
    w, h, frames = O.read_ivf(data)
    gop = frames[:12]
    # independent copies of the first GOP: truncate the 5th frame of copy 3 of 8
    streams = [list(gop) for _ in range(8)]
    streams[3][4] = streams[3][4][:7]
    bad = write_ivf(w, h, [f for s in streams for f in s])
    good = write_ivf(w, h, [f for _ in range(8) for f in gop])
    ctx = Context(w, h, max_frames=8 * 110 + 64)
    ctx.set_device_tokens(device_tokens)
    base = ctx.L.vp8gpu_frames_in_use(ctx.h)
    with pytest.raises(capi.Vp8Error):
        decode_ivf(ctx, bad, threads=4)
    ctx.sync()
    assert ctx.L.vp8gpu_frames_in_use(ctx.h) == base, "rasters leaked by the failed call"
    out, nd, ns = decode_ivf(ctx, good, threads=4)
    want = O.decode_ivf_display(write_ivf(w, h, gop))
    assert nd == 96 and out == want * 8
    assert ctx.L.vp8gpu_frames_in_use(ctx.h) == base
    ctx.close()
