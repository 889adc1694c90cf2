#!/bin/bash
# Build tuning variants of liboodb200.so (macro sweeps) into ood_in_object_detection_b200/variants/.
# usage: scripts/build_variants.sh name1:"-DX=1 -DY=2" name2:"..."
set -e
cd "$(dirname "$0")/.."
D=ood_in_object_detection_b200; mkdir -p $D/variants
SRC=$(python -c "from ood_in_object_detection_b200 import build as b; print(' '.join('$D/csrc/' + s for s in b.SOURCES))")
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O3 --shared -cudart shared $flags \
    $SRC -o $D/variants/$name.so &
done
wait
ls -la $D/variants
