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
