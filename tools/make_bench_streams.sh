#!/bin/sh
# Re-create bench_data/*.ivf (build container only): synthetic 1080p YUV420 from seeded generators,
# encoded to VP8 by the UNMODIFIED reference encoder (oracle/_ref/ref_encode, REALTIME_QUALITY,
# one Encoder per 30-frame GOP).  No >=1080p VP8 material and no external encoder exist here
# (SURVEY.md 8d).  The loop-filter level is chosen by the reference encoder through the SSIM
# restatement in oracle/ref_shim/ssim_stub.cc (parity unpinned, see that file).
set -e
cd "$(dirname "$0")/.."
make -C oracle -j8 ref >/dev/null
mkdir -p bench_data
oracle/_ref/ref_encode bench_data/synth1080p_medium_q90.ivf 1920 1080 60 30 90 1234 2 2>/dev/null
oracle/_ref/ref_encode bench_data/synth1080p_easy_q40.ivf 1920 1080 60 30 40 1234 0 2>/dev/null
oracle/_ref/ref_encode bench_data/synth720p_medium_q90.ivf 1280 720 60 30 90 4321 2 2>/dev/null   # BASELINE.json config 5 (64 x 720p)
ls -la bench_data
# short 4K clip (BASELINE.json config 4 parity case) and the reference's own answers for all clips
oracle/_ref/ref_encode bench_data/synth4k_medium_q90_8f.ivf 3840 2160 8 4 90 77 2 2>/dev/null
# feature-complete 1080p stream (SURVEY.md 8d bitstream B): written by the product's bitstream writer from
# seeded-random records; needs alfalfa_b200/libvp8gpu.so (host code only)
python tools/make_feature_stream.py bench_data/features1080p_12f.ivf 1920 1080 12 2024
python - <<'PY'
import hashlib, json, subprocess, os
out = {}
for n in sorted(os.listdir("bench_data")):
    if n.endswith(".ivf"):
        raw = subprocess.run(["oracle/_ref/ref_dump", "shown", "bench_data/" + n], capture_output=True).stdout
        out[n] = {"sha1_of_reference_decode": hashlib.sha1(raw).hexdigest(), "bytes": len(raw)}
json.dump(out, open("tests/golden/bench_clips.json", "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1))
PY
