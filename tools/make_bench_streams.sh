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
ls -la bench_data
