"""ctypes binding of the CPU oracle (oracle/libvp8oracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg, never by the
product package."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "libvp8oracle.so")


class FrameDesc(C.Structure):
    _fields_ = [("width", C.c_uint16), ("height", C.c_uint16), ("mb_cols", C.c_uint16), ("mb_rows", C.c_uint16),
                ("key_frame", C.c_uint8), ("show_frame", C.c_uint8), ("loop_filter_level", C.c_uint8),
                ("sharpness", C.c_uint8), ("pad0", C.c_uint8 * 4), ("quant", C.c_uint16 * 24),
                ("n_tokens", C.c_uint32), ("n_split", C.c_uint32), ("refresh_last", C.c_uint8),
                ("refresh_golden", C.c_uint8), ("refresh_alternate", C.c_uint8), ("copy_to_golden", C.c_uint8),
                ("copy_to_alternate", C.c_uint8), ("pad1", C.c_uint8 * 3)]


MB_DTYPE = np.dtype([("tok_off", "<u4"), ("tok_cnt", "<u2"), ("y_mode", "u1"), ("uv_mode", "u1"),
                     ("ref_frame", "u1"), ("segment_id", "u1"), ("lf_level", "u1"), ("flags", "u1"),
                     ("mv_x", "<i2"), ("mv_y", "<i2"), ("split_idx", "<u4"), ("reserved", "<u4"),
                     ("b_modes", "<u8")])
assert MB_DTYPE.itemsize == 32


class Raster(C.Structure):
    _fields_ = [("w16", C.c_int), ("h16", C.c_int), ("y", C.POINTER(C.c_uint8)), ("u", C.POINTER(C.c_uint8)),
                ("v", C.POINTER(C.c_uint8))]


class Parsed(C.Structure):
    _fields_ = [("desc", FrameDesc), ("mbs", C.c_void_p), ("tokens", C.c_void_p), ("split", C.c_void_p),
                ("mbs_cap", C.c_size_t), ("tokens_cap", C.c_size_t), ("split_cap", C.c_size_t)]


def build():
    if not os.path.exists(LIB_PATH) or any(
            os.path.getmtime(os.path.join(ORACLE_DIR, f)) > os.path.getmtime(LIB_PATH)
            for f in ("vp8_oracle.c", "vp8_oracle.h", "vp8_tables.h")):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.vp8o_decoder_new.restype = C.c_void_p
        L.vp8o_decoder_new.argtypes = [C.c_int, C.c_int]
        L.vp8o_decoder_free.argtypes = [C.c_void_p]
        L.vp8o_decoder_decode.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_int),
                                          C.POINTER(C.POINTER(Raster)), C.POINTER(Raster)]
        L.vp8o_decoder_last_parsed.restype = C.POINTER(Parsed)
        L.vp8o_decoder_last_parsed.argtypes = [C.c_void_p]
        L.vp8o_decoder_ref.restype = C.POINTER(Raster)
        L.vp8o_decoder_ref.argtypes = [C.c_void_p, C.c_int]
        L.vp8o_raster_new.restype = C.POINTER(Raster)
        L.vp8o_raster_new.argtypes = [C.c_int, C.c_int]
        L.vp8o_raster_free.argtypes = [C.POINTER(Raster)]
        L.vp8o_raster_dump_display.restype = C.c_size_t
        L.vp8o_raster_dump_display.argtypes = [C.POINTER(Raster), C.c_int, C.c_int, C.c_void_p]
        L.vp8o_time_ivf.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_uint32, C.POINTER(C.c_double),
                                    C.POINTER(C.c_uint32)]
        L.vp8o_reconstruct.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Raster),
                                       C.POINTER(Raster), C.POINTER(Raster), C.POINTER(Raster)]
        L.vp8o_loopfilter.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(Raster)]
        _lib = L
    return _lib


def read_ivf(data):
    """(width, height, [frame bytes...]) of an in-memory IVF (util/ivf.cc:36-82)."""
    assert data[:4] == b"DKIF", "not an IVF file"
    hdr_len = struct.unpack_from("<H", data, 6)[0]
    w, h = struct.unpack_from("<HH", data, 12)
    n = struct.unpack_from("<I", data, 24)[0]
    frames, pos = [], hdr_len
    for _ in range(n):
        if pos + 12 > len(data):
            break
        flen = struct.unpack_from("<I", data, pos)[0]
        frames.append(bytes(data[pos + 12:pos + 12 + flen]))
        pos += 12 + flen
    return w, h, frames


def raster_planes(r):
    """numpy copies (Y, U, V) of a vp8o_raster."""
    r = r.contents if hasattr(r, "contents") else r
    w, h = r.w16, r.h16
    y = np.ctypeslib.as_array(r.y, shape=(h, w)).copy()
    u = np.ctypeslib.as_array(r.u, shape=(h // 2, w // 2)).copy()
    v = np.ctypeslib.as_array(r.v, shape=(h // 2, w // 2)).copy()
    return y, u, v


class ParsedView:
    """numpy views of the flat records of one parsed frame (copied)."""

    def __init__(self, p):
        p = p.contents if hasattr(p, "contents") else p
        self.desc = FrameDesc.from_buffer_copy(bytes(p.desc))
        n_mbs = self.desc.mb_cols * self.desc.mb_rows
        self.mbs = np.frombuffer(C.string_at(p.mbs, n_mbs * 32), dtype=MB_DTYPE).copy()
        nt = self.desc.n_tokens
        self.tokens = np.frombuffer(C.string_at(p.tokens, nt * 4), dtype="<u4").copy() if nt else np.zeros(0, "<u4")
        ns = self.desc.n_split
        self.split = (np.frombuffer(C.string_at(p.split, ns * 64), dtype="<i2").copy().reshape(ns, 16, 2)
                      if ns else np.zeros((0, 16, 2), "<i2"))


class OracleDecoder:
    """Decoder (decoder.hh:244-300) restated on the CPU."""

    def __init__(self, width, height):
        self.L = lib()
        self.w, self.h = width, height
        self.d = self.L.vp8o_decoder_new(width, height)
        self._pre = self.L.vp8o_raster_new(width, height)

    def __del__(self):
        try:
            self.L.vp8o_decoder_free(self.d)
            self.L.vp8o_raster_free(self._pre)
        except Exception:
            pass

    def decode(self, frame, want_pre_lf=False, want_planes=True):
        """returns dict(shown, planes=(Y,U,V) after loop filter, pre=(Y,U,V) before, display=bytes)"""
        shown = C.c_int(0)
        out = C.POINTER(Raster)()
        rc = self.L.vp8o_decoder_decode(self.d, frame, len(frame), C.byref(shown), C.byref(out),
                                        self._pre if want_pre_lf else None)
        if rc != 0:
            raise ValueError("oracle decode failed: %d" % rc)
        res = {"shown": bool(shown.value), "raster": out}
        if want_planes:
            res["planes"] = raster_planes(out)
        if want_pre_lf:
            res["pre"] = raster_planes(self._pre)
        return res

    def display_bytes(self, raster):
        n = self.w * self.h + 2 * ((self.w + 1) // 2) * ((self.h + 1) // 2)
        buf = (C.c_uint8 * n)()
        got = self.L.vp8o_raster_dump_display(raster, self.w, self.h, buf)
        assert got == n
        return bytes(buf)

    def parsed(self):
        return ParsedView(self.L.vp8o_decoder_last_parsed(self.d))


def decode_ivf_display(data):
    """Every shown frame's display rectangle, concatenated -- the byte stream the reference's
    tests/decode-to-stdout.cc writes (FilePlayer: start at the first key frame)."""
    w, h, frames = read_ivf(data)
    dec = OracleDecoder(w, h)
    out = []
    started = False
    for f in frames:
        if not started and (len(f) < 1 or (f[0] & 1)):
            continue
        started = True
        r = dec.decode(f, want_planes=False)
        if r["shown"]:
            out.append(dec.display_bytes(r["raster"]))
    return b"".join(out)
