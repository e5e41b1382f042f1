# tuning sweep for the device-token pipeline of vp8gpu_decode_ivf (diagnostic; run on the GPU box)
run() { echo "== slots=$1 chunk=$2 threads=$3 $4 $5 $6"; VP8GPU_TOK_SLOTS=$1 VP8GPU_TOK_CHUNK=$2 timeout 200 python bench.py --steps 3 --warmup 1 --gop-instances 2 --no-encode --no-cpu-baseline --host-stats --threads $3 $4 $5 $6 2>&1 | grep -o "host stats.*\|\"e2e\": {\"value\": [0-9.]*"; }
run 60 15 64
run 60 15 64 --no-output
run 60 15 32
run 60 15 64 --replicas 128
