#!/usr/bin/env bash
# Kernel experiments: builds libphovo_b200 into build/variants/<name>.so (git-ignored, travels to the GPU box)
# with extra nvcc flags and, optionally, an alternative kernels_batch.cu; tools/run_variants.sh then times each
# variant with the same bench command on the box.   usage: build_variant.sh <name> [alt_kernels_batch.cu] [-- nvcc flags]
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
CS="$ROOT/photoconsistency-visual-odometry_b200/csrc"
NAME=$1; shift
ALT=""
if [ $# -gt 0 ] && [ "$1" != "--" ]; then ALT=$1; shift; fi
if [ $# -gt 0 ] && [ "$1" == "--" ]; then shift; fi
OUT="$ROOT/build/variants"; OBJ="$OUT/obj_$NAME"
mkdir -p "$OBJ"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -I "$CS" "$@")
for f in phovo_api.cu phovo_batch.cu kernels_pyramid.cu kernels_align.cu yaml_config.cpp; do
  # the other translation units do not change between variants: reuse the product objects when they are current
  if [ -f "$CS/$f.o" ] && [ "$CS/$f.o" -nt "$CS/$f" ] && [ $# -eq 0 ]; then cp "$CS/$f.o" "$OBJ/$f.o"; else nvcc "${FLAGS[@]}" -c "$CS/$f" -o "$OBJ/$f.o"; fi
done
SRC="$CS/kernels_batch.cu"; [ -n "$ALT" ] && SRC="$ALT"
nvcc "${FLAGS[@]}" -Xptxas -v -c "$SRC" -o "$OBJ/kernels_batch.cu.o" 2> "$OUT/$NAME.ptxas.log"
nvcc -shared -o "$OUT/$NAME.so" "$OBJ"/*.o -gencode arch=compute_100a,code=sm_100a -lcudart
grep -A1 "Lb0E" "$OUT/$NAME.ptxas.log" | grep -E "spill" | head -4
echo "built $OUT/$NAME.so"
