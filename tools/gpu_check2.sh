#!/bin/bash
# GPU-box script: new boundary tests, then A/B/C of wavefront variants and pipeline knobs.
TAG=${1:-x}
O=gpurun_out
mkdir -p $O
( VP8GPU_WAVEFRONT=ll timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_feature_stream.py -m gpu -x -q 2>&1 | tail -6 ) > $O/${TAG}_parity_ll.log 2>&1
( timeout 600 python -m pytest tests/test_gpu_cxx_host.py tests/test_ref_flatten.py tests/test_state_format.py tests/test_gpu_encoder.py -m gpu -q 2>&1 | tail -25 ) > $O/${TAG}_newtests.log 2>&1
tail -4 $O/${TAG}_parity_ll.log $O/${TAG}_newtests.log
run() {  # name, env..., -- args
  local name=$1; shift
  env "$@" timeout 200 python bench.py --no-encode --no-cpu-baseline --steps 3 --warmup 1 $EXTRA > $O/${TAG}_bench_${name}.json 2> $O/${TAG}_bench_${name}.err
}
EXTRA="" run default X=1
EXTRA="" run ll VP8GPU_WAVEFRONT=ll
EXTRA="" run legacy VP8GPU_WAVEFRONT=legacy
EXTRA="--threads 128" run t128 X=1
EXTRA="" run disp2 VP8GPU_DISPATCHERS=2
EXTRA="--gop-instances 128" run g128 X=1
python - <<PY
import json
for k in ("default","ll","legacy","t128","disp2","g128"):
    try:
        d=json.loads(open("$O/${TAG}_bench_%s.json"%k).read().strip().splitlines()[-1])
        r=d["roofline"]
        print("%-8s value %6.0f e2e %6.0f resident %6.0f ms/step %s" % (k, d["value"], d["e2e"]["value"], r["resident_value"], {a: round(b,1) for a,b in r["kernel_ms_per_step"].items()}))
    except Exception as e:
        print(k, "no bench line:", e)
PY
VP8GPU_WAVEFRONT=ll timeout 120 python tools/phase_profile.py --g 64 > $O/${TAG}_phase_ll.log 2>&1
grep -A9 "k_loopfilter:" $O/${TAG}_phase_ll.log | head -22
