"""YUV4MPEG2 ingest for the encoder path: the host-side mirror of YUV4MPEGReader
(/root/reference/src/input/yuv4mpeg.cc:158-300).  Same header grammar and error behaviour: magic
"YUV4MPEG2", W / H / F / I / A / C / X tokens, only the 4:2:0 colour spaces, every frame introduced by a
"FRAME" line, planar Y (w x h), U, V (w/2 x h/2 each, as the reference reads them).  Frames are returned as
display-size numpy planes; the macroblock-aligned raster with its replicated right / bottom edges
(edge_extend, yuv4mpeg.cc:231-271) is produced where the reference's reader produces it -- on the way into the
raster, which here is the device upload inside vp8gpu_encoder_encode_* ; `edge_extend` below is the numpy
restatement of that step for tests and tools."""
import numpy as np


class Y4MError(RuntimeError):
    """the reference throws std::runtime_error for every malformed input"""


def _fraction(text):
    if ":" not in text:
        raise Y4MError("invalid fraction")
    a, b = text.split(":", 1)
    return int(a), int(b)


class Y4MReader:
    def __init__(self, data):
        self.data = memoryview(data)
        end = bytes(self.data[:4096]).find(b"\n")
        if end < 0:
            raise Y4MError("invalid yuv4mpeg2 magic code")
        tokens = bytes(self.data[:end]).decode("ascii", "replace").split()
        if not tokens or tokens[0] != "YUV4MPEG2":
            raise Y4MError("invalid yuv4mpeg2 magic code")
        self.width = self.height = 0
        self.fps = (0, 0)
        self.interlacing = "p"
        self.aspect = (0, 0)
        for tok in tokens[1:]:
            kind, rest = tok[0], tok[1:]
            if kind == "W":
                self.width = int(rest)
            elif kind == "H":
                self.height = int(rest)
            elif kind == "F":
                self.fps = _fraction(rest)
            elif kind == "I":
                if not rest or rest[0] not in "ptbm":
                    raise Y4MError("invalid interlacing mode")
                self.interlacing = rest[0]
            elif kind == "A":
                self.aspect = _fraction(rest)
            elif kind == "C":
                if not tok.startswith("C420"):
                    raise Y4MError("only yuv420 color space is supported")
            elif kind == "X":
                pass
            else:
                raise Y4MError("invalid yuv4mpeg2 input format")
        if self.width == 0 or self.height == 0:
            raise Y4MError("width or height missing")
        self.pos = end + 1

    def get_next_frame(self):
        """YUV4MPEGReader::get_next_frame: (y, u, v) or None at the end of the input"""
        if self.pos >= len(self.data):
            return None
        end = bytes(self.data[self.pos:self.pos + 256]).find(b"\n")
        if end < 0 or not bytes(self.data[self.pos:self.pos + end]).startswith(b"FRAME"):
            raise Y4MError("invalid yuv4mpeg2 input format")
        p = self.pos + end + 1
        w, h = self.width, self.height
        cw, ch = w // 2, h // 2
        need = w * h + 2 * cw * ch
        if p + need > len(self.data):
            raise Y4MError("unexpected end of file")
        buf = np.frombuffer(self.data, np.uint8, need, p)
        y = buf[:w * h].reshape(h, w)
        u = buf[w * h:w * h + cw * ch].reshape(ch, cw)
        v = buf[w * h + cw * ch:].reshape(ch, cw)
        self.pos = p + need
        if (w | h) & 1:
            # the raster's chroma planes are (w + 1) / 2 x (h + 1) / 2: replicate into the missing row / column
            u = np.pad(u, ((0, (h + 1) // 2 - ch), (0, (w + 1) // 2 - cw)), mode="edge")
            v = np.pad(v, ((0, (h + 1) // 2 - ch), (0, (w + 1) // 2 - cw)), mode="edge")
        return y, u, v

    def __iter__(self):
        while True:
            f = self.get_next_frame()
            if f is None:
                return
            yield f


def write_y4m(width, height, frames, fps=(30, 1)):
    """YUV4MPEGHeader::to_string (yuv4mpeg.cc:84-124) + FRAME records"""
    out = bytearray(("YUV4MPEG2 W%d H%d F%d:%d Ip A1:1 C420 XYSCSS=420\n" % (width, height, fps[0], fps[1])).encode())
    for y, u, v in frames:
        out += b"FRAME\n"
        out += np.ascontiguousarray(y[:height, :width]).tobytes()
        out += np.ascontiguousarray(u[:height // 2, :width // 2]).tobytes()
        out += np.ascontiguousarray(v[:height // 2, :width // 2]).tobytes()
    return bytes(out)


def edge_extend(plane, aligned_width, aligned_height):
    """edge_extend_component (yuv4mpeg.cc:231-263): right, bottom, lower-right quadrant by replication"""
    h, w = plane.shape
    return np.pad(plane, ((0, aligned_height - h), (0, aligned_width - w)), mode="edge")
