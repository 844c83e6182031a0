#!/bin/bash
# Build the kernels of a git revision (default HEAD) as tools/_build/var_<name>/libflic_b200.so so that
# tools/variants.py run times them beside the working tree's build on the same box.
#   tools/build_base_variant.sh [rev] [name]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REV=${1:-HEAD}; NAME=${2:-base}
PKG="$ROOT/finalproject-losslessimagecompression_b200"
SRC="$ROOT/tools/_build/src_$NAME"; OUT="$ROOT/tools/_build/var_$NAME"
rm -rf "$SRC"; mkdir -p "$SRC" "$OUT"
git -C "$ROOT" archive "$REV" finalproject-losslessimagecompression_b200/csrc include | tar -x -C "$SRC"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=false -Xcompiler -fPIC"
OBJS=""
for f in "$SRC"/finalproject-losslessimagecompression_b200/csrc/*.cu; do
  o="$OUT/$(basename "${f%.cu}").o"
  nvcc $FLAGS -I "$SRC/finalproject-losslessimagecompression_b200/csrc" -I "$SRC/include" -c "$f" -o "$o" &
  OBJS="$OBJS $o"
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -o "$OUT/libflic_b200.so" $OBJS
echo "built $OUT/libflic_b200.so from $REV"
