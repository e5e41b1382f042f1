# tuning sweep for the device-token pipeline of vp8gpu_decode_ivf (diagnostic; run on the GPU box)
run() { echo "== slots=$1 chunk=$2 threads=$3"; VP8GPU_TOK_SLOTS=$1 VP8GPU_TOK_CHUNK=$2 timeout 200 python bench.py --steps 2 --warmup 1 --gop-instances 2 --no-encode --no-cpu-baseline --host-stats --threads $3 2>&1 | grep -o "host stats.*\|\"e2e\": {\"value\": [0-9.]*"; }
run 60 15 64
run 60 30 64
run 32 8 64
run 60 15 96
run 60 15 48
