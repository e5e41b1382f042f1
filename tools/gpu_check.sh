#!/bin/bash
# GPU-box script: parity of the current kernels, then A/B numbers against the round-1 wavefront kernels.
# usage (through gpurun): bash tools/gpu_check.sh TAG
TAG=${1:-x}
O=gpurun_out
mkdir -p $O
( timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15 ) > $O/${TAG}_parity.log 2>&1
( timeout 300 python -m pytest tests/test_feature_stream.py tests/test_gpu_tokens.py -m gpu -x -q 2>&1 | tail -15 ) > $O/${TAG}_features.log 2>&1
tail -3 $O/${TAG}_parity.log $O/${TAG}_features.log
if grep -q "failed\|error\|Error" $O/${TAG}_parity.log || ! grep -q "passed" $O/${TAG}_parity.log; then
  echo "PARITY FAILED: diagnostics"
  timeout 200 python tools/gpu_diag.py --max-frames 3 > $O/${TAG}_diag.log 2>&1; tail -30 $O/${TAG}_diag.log
fi
timeout 120 python tools/phase_profile.py --g 64 > $O/${TAG}_phase_ll.log 2>&1
VP8GPU_WAVEFRONT=legacy timeout 120 python tools/phase_profile.py --g 64 > $O/${TAG}_phase_legacy.log 2>&1
timeout 240 python bench.py --no-encode --no-cpu-baseline > $O/${TAG}_bench_ll.json 2> $O/${TAG}_bench_ll.err
VP8GPU_WAVEFRONT=legacy timeout 240 python bench.py --no-encode --no-cpu-baseline > $O/${TAG}_bench_legacy.json 2> $O/${TAG}_bench_legacy.err
python - <<PY
import json
for k in ("ll","legacy"):
    try:
        d=json.loads(open("$O/${TAG}_bench_%s.json"%k).read().strip().splitlines()[-1])
        r=d["roofline"]
        print(k, "value %.0f e2e %.0f resident %.0f ms/step %s" % (d["value"], d["e2e"]["value"], r["resident_value"], r["kernel_ms_per_step"]))
    except Exception as e:
        print(k, "no bench line:", e)
PY
grep -A12 "frame 30\|frame 33" $O/${TAG}_phase_ll.log | head -60
