"""BASELINE.json configs[0] / SURVEY.md 8(d) config 1: the 320x240 all-key-frame vector
45502fe01a62... goes Chunk -> front end of the PRODUCT through the C ABI (vp8gpu_parse_frame =
decompress_frame + parse_frame<KeyFrame>) -> the seam records of include/vp8gpu.h -> the CPU oracle's
restatement of Frame::decode + Frame::loopfilter behind that seam -> decode-to-stdout dump.  Pass = the
SHA-1 of the dump is the file name (the reference's own golden check, tests/decoding.test:14-15).
No GPU: this pins the boundary's data formats on the CPU; the same seam is what the CUDA back end
consumes (vp8gpu_decode_parsed), which the -m gpu tests compare with the oracle."""
import ctypes as C
import hashlib
import os

import oracle_lib as O
from alfalfa_b200 import capi
from conftest import GOLDEN_DIR

NAME = "45502fe01a62b82d498b83dc50824741402436db"


def test_key_frames_through_the_boundary_with_the_oracle_behind_the_seam():
    L, OL = capi.lib(), O.lib()
    w, h, frames = O.read_ivf(open(os.path.join(GOLDEN_DIR, NAME), "rb").read())
    assert (w, h, len(frames)) == (320, 240, 30)
    st, pf = C.c_void_p(), C.c_void_p()
    capi.check(L.vp8gpu_state_create(w, h, C.byref(st)))
    capi.check(L.vp8gpu_parsed_create(C.byref(pf)))
    out = OL.vp8o_raster_new(w, h)
    n = w * h + 2 * ((w + 1) // 2) * ((h + 1) // 2)
    buf = (C.c_uint8 * n)()
    sha = hashlib.sha1()
    for f in frames:
        assert not (f[0] & 1), "every frame of this vector is a key frame"
        capi.check(L.vp8gpu_parse_frame(st, f, len(f), pf), None, "parse_frame")
        desc = L.vp8gpu_parsed_desc(pf)
        mbs, tok, sp = L.vp8gpu_parsed_mbs(pf), L.vp8gpu_parsed_tokens(pf), L.vp8gpu_parsed_split(pf)
        # key frames predict from nothing: no references behind the seam
        OL.vp8o_reconstruct(desc, mbs, tok, sp, None, None, None, out)
        OL.vp8o_loopfilter(desc, mbs, out)
        assert OL.vp8o_raster_dump_display(out, w, h, buf) == n
        if desc.contents.show_frame:
            sha.update(bytes(buf))
    OL.vp8o_raster_free(out)
    L.vp8gpu_parsed_destroy(pf)
    L.vp8gpu_state_destroy(st)
    assert sha.hexdigest() == NAME
