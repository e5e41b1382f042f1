#!/usr/bin/env python3
"""GPU-box diagnostic: the sampled size estimates (Encoder::estimate_frame_size) the target-size search sees for
frame 1 of the bench's encode clip, next to the reference's (oracle/_ref/ref_encode with REF_EST_FRAME)."""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from alfalfa_b200 import Context, Encoder  # noqa: E402

w, h = 1920, 1080
target = int(sys.argv[1]) if len(sys.argv) > 1 else 45000
src = [bench.synth_1080p(t) for t in range(2)]
ctx = Context(w, h, max_frames=32)
enc = Encoder(ctx)
blob, qi = enc.encode_with_target_size(*src[0], target)
print("frame 0: %d bytes, qi %d" % (len(blob), qi))
out_dir = os.path.join(ROOT, "gpurun_out")
os.makedirs(out_dir, exist_ok=True)
ours = {}
for q in range(qi - 2, qi + 10):
    os.environ["VP8GPU_EST_DUMP"] = os.path.join(out_dir, "est_ours.%d" % q)
    ours[q] = enc.estimate_frame_size(*src[1], q)
os.environ.pop("VP8GPU_EST_DUMP")
open(os.path.join(out_dir, "est_frame0.bin"), "wb").write(blob)
ref = {}
tool = os.path.join(ROOT, "oracle", "_ref", "ref_encode")
if os.path.exists(tool):
    with tempfile.TemporaryDirectory() as d:
        raw = os.path.join(d, "src.yuv")
        with open(raw, "wb") as f:
            for t in range(2):
                for p in src[t]:
                    f.write(p.tobytes())
        env = dict(os.environ, REF_RAW=raw, REF_TARGET=str(target), REF_EST_FRAME="1", REF_EST_LO=str(qi - 2), REF_EST_HI=str(qi + 9),
                   REF_EST_DUMP=os.path.join(out_dir, "est_ref"))
        r = subprocess.run([tool, os.path.join(d, "o.ivf"), str(w), str(h), "2", "1000", "0"], env=env, capture_output=True, text=True)
        for line in r.stderr.splitlines():
            if line.startswith("estimate"):
                _, _, _, _, q, v = line.split()
                ref[int(q)] = int(v)
for q in sorted(ours):
    print("qi %3d: ours %7d reference %7s %s" % (q, ours[q], ref.get(q), "" if ref.get(q) == ours[q] else "<-- differs"))
blob1, qi1 = enc.encode_with_target_size(*src[1], target)
print("frame 1: %d bytes, qi %d" % (len(blob1), qi1))
del enc
ctx.close()
