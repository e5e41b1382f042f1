#!/bin/sh
# Build libvp8gpu.so (CUDA kernels for sm_100a + host library) in-tree.
# usage: alfalfa_b200/csrc/build.sh   -> alfalfa_b200/libvp8gpu.so
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wextra,-pthread"
mkdir -p build
$NVCC $ARCH $FLAGS -Xptxas -v -c kernels.cu -o build/kernels.o 2> build/ptxas_kernels.log || { cat build/ptxas_kernels.log; exit 1; }
$NVCC $ARCH $FLAGS -Xptxas -v -c tokens.cu -o build/tokens.o 2> build/ptxas_tokens.log || { cat build/ptxas_tokens.log; exit 1; }
$NVCC $ARCH $FLAGS -c engine.cu -o build/engine.o
$NVCC $FLAGS -x cu $ARCH -c capi.cc -o build/capi.o
$NVCC $ARCH $FLAGS -c encoder.cu -o build/encoder.o
$NVCC $FLAGS -x cu $ARCH -c comm.cc -o build/comm.o
g++ -O3 -std=c++17 -fPIC -Wall -Wextra -c parser.cc -o build/parser.o
g++ -O3 -std=c++17 -fPIC -Wall -Wextra -pthread -c serializer.cc -o build/serializer.o
g++ -O3 -std=c++17 -fPIC -Wall -Wextra -c enc_costs.cc -o build/enc_costs.o
$NVCC $ARCH -shared -o ../libvp8gpu.so build/kernels.o build/tokens.o build/engine.o build/capi.o build/encoder.o build/comm.o build/parser.o build/serializer.o build/enc_costs.o -Xcompiler -pthread -ldl
echo "built $(cd .. && pwd)/libvp8gpu.so"
# optional: variant WITH a tensormap acquire fence before every TMA copy (timing experiment, tools only)
if [ "$1" = "fence" ]; then
  $NVCC $ARCH $FLAGS -DVP8_TMAP_FENCE -c kernels.cu -o build/kernels_fence.o
  $NVCC $ARCH -shared -o ../libvp8gpu_fence.so build/kernels_fence.o build/tokens.o build/engine.o build/capi.o build/encoder.o build/comm.o build/parser.o build/serializer.o build/enc_costs.o -Xcompiler -pthread -ldl
  echo "built libvp8gpu_fence.so"
fi
# optional: phase-profiling variant of the library (tools/phase_profile.py)
if [ "$1" = "prof" ]; then
  $NVCC $ARCH $FLAGS -DVP8_PROFILE -c kernels.cu -o build/kernels_prof.o
  $NVCC $ARCH -shared -o ../libvp8gpu_prof.so build/kernels_prof.o build/tokens.o build/engine.o build/capi.o build/encoder.o build/comm.o build/parser.o build/serializer.o build/enc_costs.o -Xcompiler -pthread -ldl
  echo "built libvp8gpu_prof.so"
fi
