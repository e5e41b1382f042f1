# tuning sweep for the device-token pipeline of vp8gpu_decode_ivf (diagnostic; run on the GPU box)
run() { echo "== dispatchers=$1 threads=$2 $3 $4"; VP8GPU_DISPATCHERS=$1 timeout 200 python bench.py --steps 3 --warmup 1 --gop-instances 2 --no-encode --no-cpu-baseline --host-stats --threads $2 $3 $4 2>&1 | grep -o "dispatcher: .*\|\"e2e\": {\"value\": [0-9.]*"; }
run 1 64
run 2 64
run 4 64
run 1 96
run 2 96
