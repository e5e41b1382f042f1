#!/bin/sh
# TEST INFRASTRUCTURE: build tests/simt/_build/libvp8gpu_simt.so -- the product's sources compiled with g++ against
# the SIMT emulator (include/cuda_runtime.h).  Never loaded by the product; only tests/test_simt_*.py use it.
set -e
cd "$(dirname "$0")"
SRC=../../alfalfa_b200/csrc
mkdir -p _build
CXX="g++ -O2 -g -std=c++17 -fPIC -pthread -DVP8GPU_SIMT_EMUL -Iinclude -Wall -Wno-unknown-pragmas -Wno-unused-function -Wno-unused-variable"
for f in kernels.cu tokens.cu engine.cu encoder.cu capi.cc comm.cc; do
  $CXX -x c++ -c $SRC/$f -o _build/${f%.*}.o &
done
for f in parser.cc serializer.cc enc_costs.cc; do
  $CXX -c $SRC/$f -o _build/${f%.*}.o &
done
$CXX -c simt_runtime.cc -o _build/simt_runtime.o &
wait
g++ -shared -o _build/libvp8gpu_simt.so _build/kernels.o _build/tokens.o _build/engine.o _build/encoder.o _build/capi.o _build/comm.o \
  _build/parser.o _build/serializer.o _build/enc_costs.o _build/simt_runtime.o -pthread -ldl
echo "built $(pwd)/_build/libvp8gpu_simt.so"
