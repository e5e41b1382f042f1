# tuning sweep for the device-token pipeline of vp8gpu_decode_ivf (diagnostic; run on the GPU box)
run() { echo "== warps=$1 slots=$2 chunk=$3 threads=$4 $5 $6"; VP8GPU_TOK_WARPS=$1 VP8GPU_TOK_SLOTS=$2 VP8GPU_TOK_CHUNK=$3 timeout 200 python bench.py --steps 3 --warmup 1 --gop-instances 2 --no-encode --no-cpu-baseline --host-stats --threads $4 $5 $6 2>&1 | grep -o "\[trace\].*\|step wall [0-9.]*\|\"e2e\": {\"value\": [0-9.]*"; }
run 32 60 30 64
run 1 60 15 64
run 32 60 30 64 --replicas 128
run 32 60 30 96
