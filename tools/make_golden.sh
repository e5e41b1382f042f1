#!/bin/sh
# Re-create tests/golden/vectors from the reference tree (build container only) and verify each
# vector against the unmodified reference decoder compiled by oracle/Makefile.
set -e
cd "$(dirname "$0")/.."
make -C oracle ref >/dev/null
mkdir -p tests/golden/vectors
for f in /root/reference/alfalfa_test_vectors/*; do
  b=$(basename "$f"); [ "$b" = README ] && continue
  s=$(oracle/_ref/ref_dump shown "$f" | sha1sum | cut -d' ' -f1)
  [ "$s" = "$b" ] || { echo "reference disagrees with golden name: $b"; exit 1; }
  cp "$f" tests/golden/vectors/
done
echo "golden vectors: $(ls tests/golden/vectors | wc -l) verified"
