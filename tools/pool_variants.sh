#!/usr/bin/env bash
# On the GPU box: tools/bench_pool.py (wave path throughput) for every build/variants/*.so.  usage: pool_variants.sh [pairs]
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
LIB="$ROOT/photoconsistency-visual-odometry_b200/libphovo_b200.so"
cp "$LIB" /tmp/libphovo_b200.product.so
for v in "$ROOT"/build/variants/*.so; do
  cp "$v" "$LIB"
  echo "$(basename "$v" .so) $(python "$ROOT/tools/bench_pool.py" 4 "${1:-2048}" 2>&1 | tail -1)"
done | tee "$ROOT/gpurun_out/pool_variants.txt"
cp /tmp/libphovo_b200.product.so "$LIB"
