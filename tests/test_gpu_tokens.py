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


@pytest.mark.parametrize("variant", ["8", "32"])
def test_other_launch_shapes_of_the_token_kernel(variant):
    """VP8GPU_TOK_WARPS: 8 frames per CTA, and k_tokens_lockstep (one lane per frame, the decoder as a
    one-decision-per-iteration state machine).  The knob is read once per process, hence the subprocess."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import hashlib, os, sys
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import oracle_lib as O
from alfalfa_b200 import Context, Decoder, decode_ivf
for name in ("0b546dad90ddefea5085c7751b5fa2f117630b1c", "ff2941dde20090835032c32c0644b6d401610c57"):
    data = open(os.path.join(%r, "tests", "golden", "vectors", name), "rb").read()
    w, h, frames = O.read_ivf(data)
    ctx = Context(w, h, max_frames=64)
    host, dev = Decoder(ctx), Decoder(ctx)
    for f in frames[:12]:
        a, b = host.parse_frame(f), dev.parse_frame_device(f)
        assert bytes(a.desc) == bytes(b.desc)
        for x, y in zip(a.arrays(), b.arrays()):
            assert x.tobytes() == y.tobytes()
    out, _, _ = decode_ivf(ctx, data, threads=2)
    assert hashlib.sha1(out).hexdigest() == name
    del host, dev
    ctx.close()
print("ok")
''' % (root, root, root)
    env = dict(os.environ, VP8GPU_TOK_WARPS=variant)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout[-500:] + out.stderr[-2000:]
