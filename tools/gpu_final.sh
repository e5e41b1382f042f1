#!/bin/bash
# GPU-box script: the round's final evidence.  usage (through gpurun): bash tools/gpu_final.sh
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > $O/r2_final_pytest.log 2>&1
tail -3 $O/r2_final_pytest.log
timeout 400 python bench.py > $O/r2_final_bench.json 2> $O/r2_final_bench.err
timeout 300 python bench.py --impl reference > $O/r2_final_bench_reference.json 2> $O/r2_final_bench_reference.err
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r2_final_launches.csv python bench.py --steps 2 --warmup 1 --no-encode --no-cpu-baseline > $O/r2_final_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_inter|k_intra|k_loopfilter" -c 9 -f -o $O/r2_final_full python bench.py --steps 1 --warmup 1 --no-encode --no-cpu-baseline --gop-instances 64 > $O/r2_final_ncu_full.log 2>&1
timeout 300 compute-sanitizer --tool memcheck --launch-timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_final_memcheck.log 2>&1
timeout 300 compute-sanitizer --tool racecheck --launch-timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_final_racecheck.log 2>&1
# the encoder's timeline: both search modes, then a launch list of the batched one (what the round-2 changes need next)
timeout 60 python tools/enc_search_ab.py 8 45000 > $O/r2_final_enc_ab_on.json 2> $O/r2_final_enc_ab_on.err
VP8GPU_ENC_SPECULATE=0 timeout 60 python tools/enc_search_ab.py 8 45000 > $O/r2_final_enc_ab_off.json 2> $O/r2_final_enc_ab_off.err
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2_final_enc_launches.csv python tools/enc_search_ab.py 4 45000 > $O/r2_final_enc_ncu.log 2>&1
timeout 120 python tools/reencode_bench.py --frames 8 > $O/r2_final_reencode_bench.json 2> $O/r2_final_reencode_bench.err
for wl in 4k 720p features; do
  timeout 240 python bench.py --workload $wl --no-encode --no-cpu-baseline --steps 3 --warmup 1 > $O/r2_final_bench_$wl.json 2> $O/r2_final_bench_$wl.err
done
python - <<PY
import json
for k in ("", "_reference", "_4k", "_720p", "_features"):
    try:
        d=json.loads([l for l in open("$O/r2_final_bench%s.json"%k) if l.startswith("{")][-1])
        print(k or "default", "value %.0f e2e %.0f" % (d["value"], d["e2e"]["value"]), d.get("roofline",{}).get("kernel_ms_per_step"), (d.get("encode") or {}).get("fps"), ((d.get("encode") or {}).get("reference") or {}).get("identical_frames"))
    except Exception as e:
        print(k, "no line:", e)
PY
grep -h "ERROR SUMMARY\|RACECHECK SUMMARY" $O/r2_final_memcheck.log $O/r2_final_racecheck.log
