#!/usr/bin/env python3
"""GPU-box diagnostic: decode golden vectors frame by frame through the seam
(vp8gpu_decode_parsed) with the oracle's reference rasters uploaded for every frame, so that a
wrong macroblock is reported where it first appears instead of after it has propagated.
Checks the raster before the loop filter (loop_filter_level forced to 0) and after it.

usage: tools/gpu_diag.py [--max-frames N] [--vectors a,b,...] [--verbose]
"""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
from alfalfa_b200 import Context, capi  # noqa: E402

MODE = ["DC", "V", "H", "TM", "B_PRED", "NEAREST", "NEAR", "ZERO", "NEW", "SPLIT"]


def first_bad_mbs(got, want, mbs, cols, limit=4):
    out = []
    for pi, (g, w, sz) in enumerate(zip(got, want, (16, 8, 8))):
        d = g != w
        if not d.any():
            continue
        ys, xs = np.nonzero(d)
        seen = set()
        for y, x in zip(ys, xs):
            mb = (int(y) // sz, int(x) // sz)
            if mb in seen:
                continue
            seen.add(mb)
            m = mbs[mb[0] * cols + mb[1]]
            out.append("plane %s MB(col %d,row %d) px(%d,%d) got %d want %d | y_mode %s uv %d ref %d mv (%d,%d) ntok %d lf %d flags %d"
                       % ("YUV"[pi], mb[1], mb[0], x, y, g[y, x], w[y, x], MODE[m["y_mode"]], m["uv_mode"],
                          m["ref_frame"], m["mv_x"], m["mv_y"], m["tok_cnt"], m["lf_level"], m["flags"]))
            if len(out) >= limit:
                return out, int(sum((gg != ww).sum() for gg, ww in zip(got, want)))
    return out, int(sum((gg != ww).sum() for gg, ww in zip(got, want)))


def run_vector(path, max_frames, verbose):
    data = open(path, "rb").read()
    w, h, frames = O.read_ivf(data)
    ctx = Context(w, h, max_frames=16)
    L = ctx.L
    od = O.OracleDecoder(w, h)
    ref_h = [ctx.alloc_frame() for _ in range(3)]
    out_h = ctx.alloc_frame()
    bad_pre = bad_post = n = 0
    started = False
    for fi, f in enumerate(frames):
        if not started and (f[0] & 1):
            continue
        started = True
        if n >= max_frames:
            break
        refs = [O.raster_planes(od.L.vp8o_decoder_ref(od.d, k)) for k in range(3)]
        r = od.decode(f, want_pre_lf=True)
        p = od.parsed()
        for k in range(3):
            ref_h[k].upload(*refs[k])
        ids = (C.c_int32 * 3)(*[x.id for x in ref_h])
        mbs = np.ascontiguousarray(p.mbs)
        tok = np.ascontiguousarray(p.tokens)
        sp = np.ascontiguousarray(p.split)
        for stage, want in (("pre-lf", r["pre"]), ("post-lf", r["planes"])):
            desc = capi.FrameDesc.from_buffer_copy(bytes(p.desc))
            if stage == "pre-lf":
                desc.loop_filter_level = 0
            capi.check(L.vp8gpu_decode_parsed(ctx.h, 0, C.byref(desc), mbs.ctypes.data, tok.ctypes.data if tok.size else None,
                                              sp.ctypes.data if sp.size else None, ids, out_h.id), ctx.h, "decode_parsed")
            got = out_h.planes()
            if any((g != w_).any() for g, w_ in zip(got, want)):
                lines, npx = first_bad_mbs(got, want, mbs, p.desc.mb_cols)
                if stage == "pre-lf":
                    bad_pre += 1
                else:
                    bad_post += 1
                if verbose or (bad_pre + bad_post) <= 3:
                    print("  frame %d (%s) %s: %d px differ" % (fi, "key" if p.desc.key_frame else "inter", stage, npx))
                    for ln in lines:
                        print("    " + ln)
                if stage == "pre-lf":
                    break  # post-lf would differ too
        n += 1
    for x in ref_h + [out_h]:
        x.release()
    ctx.close()
    return n, bad_pre, bad_post


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-frames", type=int, default=12)
    ap.add_argument("--vectors", default="")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    d = os.path.join(ROOT, "tests", "golden", "vectors")
    names = [v for v in a.vectors.split(",") if v] or sorted(os.listdir(d))
    t0 = time.time()
    tot = [0, 0, 0]
    for name in names:
        full = [x for x in os.listdir(d) if x.startswith(name)][0]
        n, bp, bq = run_vector(os.path.join(d, full), a.max_frames, a.verbose)
        print("%s frames %d  pre-lf bad %d  post-lf bad %d" % (full[:12], n, bp, bq), flush=True)
        tot[0] += n
        tot[1] += bp
        tot[2] += bq
    print("TOTAL frames %d pre-lf bad %d post-lf bad %d  (%.1fs)" % (tot[0], tot[1], tot[2], time.time() - t0))


if __name__ == "__main__":
    main()
