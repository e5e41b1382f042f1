#!/usr/bin/env python3
"""Latency of the device-side token decoder for single frames: vp8gpu_parse_frame_device (first
partition on the host, H2D, k_tokens with ONE frame, records back) against vp8gpu_parse_frame (all on
the host).  k_tokens runs one thread per frame, so this is the per-frame latency that the pipelined
vp8gpu_decode_ivf hides by keeping hundreds of frames in flight.
usage: python tools/tokens_bench.py [ivf] [--frames N]"""
import argparse
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("ivf", nargs="?", default=os.path.join(ROOT, "bench_data", "synth1080p_medium_q90.ivf"))
    ap.add_argument("--frames", type=int, default=12)
    a = ap.parse_args()
    from alfalfa_b200 import Context, capi
    from alfalfa_b200.decoder import read_ivf
    w, h, frames = read_ivf(open(a.ivf, "rb").read())
    L = capi.lib()
    ctx = Context(w, h, max_frames=4)
    st_h, st_d, pf = C.c_void_p(), C.c_void_p(), C.c_void_p()
    capi.check(L.vp8gpu_state_create(w, h, C.byref(st_h)))
    capi.check(L.vp8gpu_state_create(w, h, C.byref(st_d)))
    capi.check(L.vp8gpu_parsed_create(C.byref(pf)))
    for i, f in enumerate(frames[:a.frames]):
        t0 = time.perf_counter()
        capi.check(L.vp8gpu_parse_frame(st_h, f, len(f), pf), ctx.h, "parse")
        t1 = time.perf_counter()
        capi.check(L.vp8gpu_parse_frame_device(ctx.h, st_d, f, len(f), pf), ctx.h, "parse_device")
        t2 = time.perf_counter()
        d = L.vp8gpu_parsed_desc(pf).contents
        print("frame %2d %6d B %6d tokens: host %.2f ms, device path %.2f ms" % (i, len(f), d.n_tokens, (t1 - t0) * 1e3, (t2 - t1) * 1e3))
    ctx.close()


if __name__ == "__main__":
    main()
