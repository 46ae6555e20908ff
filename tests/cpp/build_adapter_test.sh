#!/bin/bash
# Compiles the drop-in C++ adapter (include/CPhotoconsistencyOdometryCuda.h) against the stand-in
# cv::/Eigen types and links the app mirror against the in-tree libphovo_b200.so.  No OpenCV/Eigen here.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
PKG="$ROOT/photoconsistency-visual-odometry_b200"
mkdir -p "$HERE/_build"
g++ -std=c++11 -O2 -Wall -Wextra -I "$ROOT/include" -I "$HERE/shim" "$HERE/frame_alignment_app.cpp" \
    -o "$HERE/_build/frame_alignment_app" -L "$PKG" -lphovo_b200 -Wl,-rpath,"\$ORIGIN/../../../photoconsistency-visual-odometry_b200" \
    -L /usr/local/cuda/lib64 -Wl,-rpath,/usr/local/cuda/lib64 -lcudart
echo "built $HERE/_build/frame_alignment_app"
