#!/bin/sh
# TEST INFRASTRUCTURE: build tests/simt/_build/libvp8gpu_simt.so -- the product's sources compiled with g++ against
# the SIMT emulator (include/cuda_runtime.h).  Never loaded by the product; only tests/test_simt_*.py use it.
set -e
cd "$(dirname "$0")"
SRC=../../alfalfa_b200/csrc
mkdir -p _build
# SIMT_SANITIZE=alignment (or address, undefined ...): the same build under a compiler sanitizer, into _build/san/ --
# a misaligned uint4 / uint2 access passes unnoticed on x86 and faults on the GPU
OUTDIR=_build
SAN=""
if [ -n "$SIMT_SANITIZE" ]; then
  OUTDIR=_build/san
  SAN="-fsanitize=$SIMT_SANITIZE -fno-sanitize-recover=all -fno-omit-frame-pointer"
  mkdir -p $OUTDIR
fi
CXX="g++ -O2 -g -std=c++17 -fPIC -pthread -DVP8GPU_SIMT_EMUL -Iinclude -Wall -Wno-unknown-pragmas -Wno-unused-function -Wno-unused-variable $SAN"
PIDS=""
for f in kernels.cu tokens.cu engine.cu encoder.cu capi.cc comm.cc; do
  $CXX -x c++ -c $SRC/$f -o $OUTDIR/${f%.*}.o &
  PIDS="$PIDS $!"
done
for f in parser.cc serializer.cc enc_costs.cc; do
  $CXX -c $SRC/$f -o $OUTDIR/${f%.*}.o &
  PIDS="$PIDS $!"
done
$CXX -c simt_runtime.cc -o $OUTDIR/simt_runtime.o &
PIDS="$PIDS $!"
for p in $PIDS; do wait $p; done   # (set -e: a failed compile fails the build instead of linking a stale object)
g++ -shared $SAN -o $OUTDIR/libvp8gpu_simt.so $OUTDIR/kernels.o $OUTDIR/tokens.o $OUTDIR/engine.o $OUTDIR/encoder.o $OUTDIR/capi.o $OUTDIR/comm.o \
  $OUTDIR/parser.o $OUTDIR/serializer.o $OUTDIR/enc_costs.o $OUTDIR/simt_runtime.o -pthread -ldl
echo "built $(pwd)/$OUTDIR/libvp8gpu_simt.so"
