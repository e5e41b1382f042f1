"""YUV4MPEG2 ingest (alfalfa_b200/y4m.py = YUV4MPEGReader, input/yuv4mpeg.cc:158-300) and the IVF minihash
field (util/ivf.cc:46, ivf_writer.cc:92-99): host plumbing, no GPU."""
import numpy as np
import pytest

from alfalfa_b200 import decoder as D
from alfalfa_b200 import y4m


def _frames(w, h, n):
    rng = np.random.default_rng(7)
    return [(rng.integers(0, 256, (h, w), np.uint8), rng.integers(0, 256, ((h + 1) // 2, (w + 1) // 2), np.uint8),
             rng.integers(0, 256, ((h + 1) // 2, (w + 1) // 2), np.uint8)) for _ in range(n)]


@pytest.mark.parametrize("size", [(64, 48), (176, 144), (1920, 1080)])
def test_round_trip(size):
    w, h = size
    frames = _frames(w, h, 3)
    r = y4m.Y4MReader(y4m.write_y4m(w, h, frames))
    assert (r.width, r.height, r.fps, r.interlacing) == (w, h, (30, 1), "p")
    got = list(r)
    assert len(got) == 3
    for (y, u, v), (y2, u2, v2) in zip(frames, got):
        assert np.array_equal(y, y2) and np.array_equal(u, u2) and np.array_equal(v, v2)
    assert r.get_next_frame() is None


def test_header_errors_like_the_reference():
    with pytest.raises(y4m.Y4MError, match="magic"):
        y4m.Y4MReader(b"YUV4MPEG W16 H16\n")
    with pytest.raises(y4m.Y4MError, match="yuv420"):
        y4m.Y4MReader(b"YUV4MPEG2 W16 H16 C444\n")
    with pytest.raises(y4m.Y4MError, match="missing"):
        y4m.Y4MReader(b"YUV4MPEG2 W16 F30:1\n")
    with pytest.raises(y4m.Y4MError, match="input format"):
        y4m.Y4MReader(b"YUV4MPEG2 W16 H16 Q1\n")
    with pytest.raises(y4m.Y4MError, match="interlacing"):
        y4m.Y4MReader(b"YUV4MPEG2 W16 H16 Ix\n")
    r = y4m.Y4MReader(b"YUV4MPEG2 W16 H16 C420jpeg XYSCSS=420JPEG\nFRAMX\n" + bytes(384))
    with pytest.raises(y4m.Y4MError):
        r.get_next_frame()


def test_edge_extension_replicates_right_bottom_and_corner():
    p = np.arange(12, dtype=np.uint8).reshape(3, 4)
    e = y4m.edge_extend(p, 6, 5)
    assert e.shape == (5, 6) and np.array_equal(e[:3, :4], p)
    assert np.all(e[:3, 4:] == p[:, 3:4]) and np.all(e[3:, :4] == p[2:3, :]) and np.all(e[3:, 4:] == p[2, 3])


def test_ivf_minihash_field():
    blob = D.write_ivf(64, 48, [b"abc", b"defg"], expected_decoder_minihash=0xDEADBEEF)
    assert D.ivf_expected_decoder_minihash(blob) == 0xDEADBEEF
    w, h, frames = D.read_ivf(blob)
    assert (w, h, frames) == (64, 48, [b"abc", b"defg"])
    assert D.ivf_expected_decoder_minihash(D.write_ivf(64, 48, [])) == 0
