#!/usr/bin/env bash
# On the GPU box: single-pair Ceres-mode latency (BASELINE configs[2]) for every build/variants/*.so
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
LIB="$ROOT/photoconsistency-visual-odometry_b200/libphovo_b200.so"
cp "$LIB" /tmp/libphovo_b200.product.so
for v in "$ROOT"/build/variants/*.so; do
  cp "$v" "$LIB"
  echo "$(basename "$v" .so) $(python "$ROOT/tools/bench_latency.py" --mode ceres 2>&1 | tail -1 | cut -c1-400)"
done | tee "$ROOT/gpurun_out/ceres_variants.txt"
cp /tmp/libphovo_b200.product.so "$LIB"
