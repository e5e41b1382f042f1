"""Device-side token decoding (alfalfa_b200/csrc/tokens.cu, k_tokens) on the B200, through the C ABI:
(a) vp8gpu_parse_frame_device must return the records vp8gpu_parse_frame returns, byte for byte, on
every frame of every golden vector (Frame::parse_tokens, frame.cc:122-137; tokens.cc:50-135);
(b) a Decoder that leaves the DCT partitions to the device reproduces the golden SHA-1s.
Bit-exact (integer work)."""
import ctypes as C
import hashlib
import os

import pytest

import oracle_lib as O
from conftest import GOLDEN_DIR, golden_vectors

pytestmark = pytest.mark.gpu


def _read(name):
    return open(os.path.join(GOLDEN_DIR, name), "rb").read()


@pytest.mark.parametrize("name", golden_vectors())
def test_device_records_identical_to_host_front_end(name):
    from alfalfa_b200 import Context, Decoder
    data = _read(name)
    w, h, frames = O.read_ivf(data)
    ctx = Context(w, h, max_frames=8)
    host, dev = Decoder(ctx), Decoder(ctx)
    started = False
    limit = 20 if w * h > 500000 else 120
    for i, f in enumerate(frames[:limit]):
        if not started and (f[0] & 1):
            continue
        started = True
        a, b = host.parse_frame(f), dev.parse_frame_device(f)
        assert bytes(a.desc) == bytes(b.desc), "frame %d desc" % i
        for x, y, what in zip(a.arrays(), b.arrays(), ("mbs", "tokens", "split")):
            assert x.tobytes() == y.tobytes(), "frame %d %s" % (i, what)
    assert host.get_state() == dev.get_state()
    del host, dev
    ctx.close()


@pytest.mark.parametrize("name", golden_vectors())
def test_decoder_with_device_tokens_reproduces_golden_sha1(name):
    from alfalfa_b200 import Context, Decoder
    data = _read(name)
    w, h, frames = O.read_ivf(data)
    ctx = Context(w, h, max_frames=16)
    dec = Decoder(ctx)
    dec.set_device_tokens(True)
    sha = hashlib.sha1()
    started = False
    for f in frames:
        if not started and (f[0] & 1):
            continue
        started = True
        shown, raster = dec.get_frame_output(f)
        if shown:
            sha.update(raster.display_bytes())
        raster.release()
    del dec
    ctx.close()
    assert sha.hexdigest() == name
