#!/usr/bin/env bash
# On the GPU box: time every build/variants/*.so with the same bench command (device-resident value only).
# The product library is put back at the end.   usage: run_variants.sh [bench args]
set -uo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
LIB="$ROOT/photoconsistency-visual-odometry_b200/libphovo_b200.so"
cp "$LIB" /tmp/libphovo_b200.product.so
mkdir -p "$ROOT/gpurun_out"
for v in "$ROOT"/build/variants/*.so; do
  n=$(basename "$v" .so)
  cp "$v" "$LIB"
  for rep in 1 2; do
    python "$ROOT/bench.py" --no-cpu-baseline --no-secondary "$@" 2> "$ROOT/gpurun_out/variant_$n.err" | python -c "
import sys, json
l = json.loads(sys.stdin.read())
print('$n', 'value %.0f' % l['value'], 'ms/step %.3f' % l['ms_per_step'], 'align_ms %.3f' % l['roofline']['kernel_ms'], 'clk', l['clocks']['sm_mhz'], l['clocks']['reasons'])
" || tail -3 "$ROOT/gpurun_out/variant_$n.err"
  done
done | tee "$ROOT/gpurun_out/variants.txt"
cp /tmp/libphovo_b200.product.so "$LIB"
