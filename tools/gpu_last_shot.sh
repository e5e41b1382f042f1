#!/bin/bash
# GPU-box script for the very last seconds of a round's GPU budget: the code written after the last hardware run, each
# step under its own timeout, every result written to gpurun_out/ as soon as it exists.
# usage (through gpurun): bash tools/gpu_last_shot.sh
O=gpurun_out
mkdir -p $O
( timeout 12 python tools/enc_search_ab.py 5 45000 > $O/r2zz_enc_ab_on.json 2> $O/r2zz_enc_ab_on.err; echo "on: $?"; tail -c 600 $O/r2zz_enc_ab_on.json ) &
wait
( VP8GPU_ENC_SPECULATE=0 timeout 12 python tools/enc_search_ab.py 5 45000 > $O/r2zz_enc_ab_off.json 2> $O/r2zz_enc_ab_off.err; echo "off: $?"; tail -c 600 $O/r2zz_enc_ab_off.json )
timeout 15 python tools/reencode_bench.py --frames 4 > $O/r2zz_reencode_bench.json 2> $O/r2zz_reencode_bench.err; echo "reencode: $?"; tail -c 400 $O/r2zz_reencode_bench.json
timeout 20 python tools/gpu_new_kernels_check.py $O/r2zz_new_kernels.json > $O/r2zz_new_kernels.log 2>&1; echo "checks: $?"; tail -3 $O/r2zz_new_kernels.log
