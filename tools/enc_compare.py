#!/usr/bin/env python3
"""GPU-box tool: our encoder vs the unmodified reference encoder (oracle/_ref/ref_encode) at fixed
quantiser indices on the same raw 1080p frames: bytes/frame, PSNR-Y and SSIM-Y (the x264-style SSIM
the reference uses, restated in numpy) of each encoder's own reconstruction, and our encode time.
usage: tools/enc_compare.py [--frames N] [--size WxH] [--qis 40,60,80,100]"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import oracle_lib as O  # noqa: E402
from alfalfa_b200 import Context, Encoder  # noqa: E402


def ssim_x264(a, b):
    """x264 pixel_ssim_wxh: 4x4 sums, 8x8 windows stepped by 4 (see oracle/ref_shim/ssim_stub.cc)"""
    a = a.astype(np.int64)
    b = b.astype(np.int64)
    h, w = (a.shape[0] // 4) * 4, (a.shape[1] // 4) * 4
    a, b = a[:h, :w], b[:h, :w]

    def blk(x):
        return x.reshape(h // 4, 4, w // 4, 4).sum(axis=(1, 3))
    s1, s2, ss, s12 = blk(a), blk(b), blk(a * a) + blk(b * b), blk(a * b)

    def win(x):
        return x[:-1, :-1] + x[1:, :-1] + x[:-1, 1:] + x[1:, 1:]
    s1, s2, ss, s12 = win(s1), win(s2), win(ss), win(s12)
    c1, c2 = int(.01 * .01 * 255 * 255 * 64 + .5), int(.03 * .03 * 255 * 255 * 64 * 63 + .5)
    var = ss * 64 - s1 * s1 - s2 * s2
    cov = s12 * 64 - s1 * s2
    v = ((2 * s1 * s2 + c1).astype(np.float64) * (2 * cov + c2)) / ((s1 * s1 + s2 * s2 + c1).astype(np.float64) * (var + c2))
    return float(v.mean())


def psnr(a, b):
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=6)
    ap.add_argument("--size", default="1920x1080")
    ap.add_argument("--qis", default="40,60,80,100")
    a = ap.parse_args()
    w, h = map(int, a.size.split("x"))
    src = [bench.synth_1080p(t, w, h) for t in range(a.frames)]
    ref_enc = os.path.join(ROOT, "oracle", "_ref", "ref_encode")
    rows = []
    for qi in map(int, a.qis.split(",")):
        ctx = Context(w, h, max_frames=16)
        enc = Encoder(ctx)
        sizes, ps, ss = [], [], []
        t0 = time.perf_counter()
        recs = []
        for t in range(a.frames):
            blob = enc.encode_with_quantizer(*src[t], qi)
            sizes.append(len(blob))
            r = enc.reconstruction()
            recs.append(r.planes()[0][:h, :w])
            r.release()
        dt = time.perf_counter() - t0
        for t in range(a.frames):
            ps.append(psnr(recs[t], src[t][0]))
            ss.append(ssim_x264(recs[t], src[t][0]))
        del enc
        ctx.close()
        row = {"qi": qi, "ours_bytes": sum(sizes) / a.frames, "ours_psnr": sum(ps) / a.frames, "ours_ssim": sum(ss) / a.frames,
               "ours_ms_per_frame": 1e3 * dt / a.frames, "ours_sizes": sizes}
        if os.path.exists(ref_enc):
            with tempfile.TemporaryDirectory() as d:
                raw = os.path.join(d, "s.yuv")
                with open(raw, "wb") as f:
                    for t in range(a.frames):
                        for p in src[t]:
                            f.write(p.tobytes())
                ivf = os.path.join(d, "o.ivf")
                r = subprocess.run([ref_enc, ivf, str(w), str(h), str(a.frames), "1000", str(qi)], env=dict(os.environ, REF_RAW=raw),
                                   capture_output=True, text=True)
                j = json.loads(r.stdout.strip().splitlines()[-1])
                data = open(ivf, "rb").read()
                _, _, frames = O.read_ivf(data)
                od = O.OracleDecoder(w, h)
                rs = []
                for t, f in enumerate(frames):
                    y = od.decode(f)["planes"][0][:h, :w]
                    rs.append(ssim_x264(y, src[t][0]))
                row.update({"ref_bytes": j["bytes"] / a.frames, "ref_psnr": j["psnr_y"], "ref_ssim": sum(rs) / a.frames,
                            "ref_ms_per_frame": 1e3 * j["encode_s"] / a.frames, "ref_sizes": [len(f) for f in frames]})
        rows.append(row)
        print(json.dumps(row), flush=True)
    print("\nqi   ours KB  ref KB  size ratio | ours PSNR  ref PSNR | ours SSIM  ref SSIM | ours ms  ref ms")
    for r in rows:
        if "ref_bytes" in r:
            print("%3d  %7.1f %7.1f  %9.2f | %9.2f %9.2f | %9.4f %9.4f | %7.1f %7.1f" % (
                r["qi"], r["ours_bytes"] / 1e3, r["ref_bytes"] / 1e3, r["ours_bytes"] / r["ref_bytes"], r["ours_psnr"], r["ref_psnr"],
                r["ours_ssim"], r["ref_ssim"], r["ours_ms_per_frame"], r["ref_ms_per_frame"]))


if __name__ == "__main__":
    main()
