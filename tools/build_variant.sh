#!/bin/bash
# Build an A/B variant of libdesta_b200.so with extra -D flags for ONE source file (the other objects are reused from the
# regular build).  Usage: tools/build_variant.sh NAME SOURCE.cu -DFLAG ...   ->  tools/_bin/libdesta_NAME.so
# Run it with DESTA_B200_LIB=tools/_bin/libdesta_NAME.so (see desta_b200/_lib.py).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; SRC=$2; shift 2
PKG="$ROOT/desta2.5-audio_b200"
mkdir -p "$ROOT/tools/_bin"
OBJ="$ROOT/tools/_bin/${NAME}_${SRC%.cu}.o"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I "$ROOT/include" -I "$PKG/csrc" "$@" \
  -c "$PKG/csrc/$SRC" -o "$OBJ"
OBJS=""
for f in "$PKG"/build/*.o; do
  if [ "$(basename "$f")" != "${SRC%.cu}.o" ]; then OBJS="$OBJS $f"; fi
done
nvcc -shared -o "$ROOT/tools/_bin/libdesta_${NAME}.so" $OBJS "$OBJ" -gencode arch=compute_100a,code=sm_100a -lcudart_static -ldl -lrt -lpthread
echo "$ROOT/tools/_bin/libdesta_${NAME}.so"
