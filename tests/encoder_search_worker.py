"""Child process of tests/test_gpu_encoder.py (not a test file): a short Salsify-shaped session -- a key frame and
inter frames by encode_with_target_size, then one encode_with_minimum_ssim -- whose result is printed as JSON.  The
test runs it twice, with the encoder's searches coding their candidates in one launch (default) and one by one
(VP8GPU_ENC_SPECULATE=0, read once per process), and expects the same bytes.

usage: python encoder_search_worker.py W H FRAMES TARGET"""
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    from alfalfa_b200 import Context, Encoder
    from test_gpu_encoder import synth
    w, h, n, target = (int(a) for a in sys.argv[1:5])
    ctx = Context(w, h, max_frames=32)
    enc = Encoder(ctx)
    out = []
    l0 = ctx.launch_count()
    for t in range(n):
        blob, qi = enc.encode_with_target_size(*synth(w, h, t), target if t else 4 * target)
        out.append({"qi": qi, "bytes": len(blob), "sha1": hashlib.sha1(bytes(blob)).hexdigest(), "stats": repr(enc.stats())})
    timeline = enc.timeline()  # of the last encode_with_target_size call
    blob, qi = enc.encode_with_minimum_ssim(*synth(w, h, n), 0.93)
    out.append({"qi": qi, "bytes": len(blob), "sha1": hashlib.sha1(bytes(blob)).hexdigest(), "stats": repr(enc.stats())})
    out.append({"minihash": enc.minihash()})
    launches = ctx.launch_count() - l0
    del enc
    ctx.close()
    print(json.dumps({"frames": out, "launches": launches, "timeline": timeline}))


if __name__ == "__main__":
    main()
