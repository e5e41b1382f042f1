#!/usr/bin/env python
"""encode_with_target_size at 1080p, a few frames, one JSON line: per-frame milliseconds, SHA-1 of every emitted frame,
kernel launches.  Run once with the encoder's searches in one launch each (default) and once with
VP8GPU_ENC_SPECULATE=0 (candidate by candidate; read once per process): same SHA-1s expected, the times differ.
usage: python tools/enc_search_ab.py [frames] [target_bytes]"""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from alfalfa_b200 import Context, Encoder
    from test_gpu_encoder import synth
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    target = int(sys.argv[2]) if len(sys.argv) > 2 else 45000
    w, h = (int(os.environ.get("AB_W", "1920")), int(os.environ.get("AB_H", "1080")))
    frames = [synth(w, h, t) for t in range(n)]
    ctx = Context(w, h, max_frames=32)
    enc = Encoder(ctx)
    l0 = ctx.launch_count()
    ms, sha, qis, phases = [], [], [], []
    for t in range(n):
        t0 = time.perf_counter()
        blob, qi = enc.encode_with_target_size(*frames[t], target)
        ms.append(round((time.perf_counter() - t0) * 1e3, 3))
        sha.append(hashlib.sha1(bytes(blob)).hexdigest()[:16])
        qis.append(qi)
        phases.append(enc.timeline())
    t0 = time.perf_counter()
    copy = enc.copy()
    t_copy1 = (time.perf_counter() - t0) * 1e3
    del copy
    t0 = time.perf_counter()
    copy = enc.copy()
    t_copy2 = (time.perf_counter() - t0) * 1e3
    del copy
    out = {"speculate": os.environ.get("VP8GPU_ENC_SPECULATE", "1"), "frames": n, "target": target, "ms": ms, "qi": qis, "sha1": sha,
           "inter_fps": round(1e3 * (n - 1) / sum(ms[1:]), 2) if n > 1 else None, "launches": int(ctx.launch_count() - l0),
           "encoder_copy_ms": [round(t_copy1, 3), round(t_copy2, 3)],
           "inter_frame_phases_ms": {k: round(sum(p[k] for p in phases[1:]) / max(1, n - 1), 3) for k in phases[0]}}
    del enc
    ctx.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
