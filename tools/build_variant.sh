#!/bin/bash
# Build a variant of the library under build/ for an A/B visit:  tools/build_variant.sh NAME [-DFLAG=VALUE ...]
# (CUDACOLLISION_LIB=$PWD/build/libcc_NAME.so selects it; build/ is git-ignored but travels to the GPU box)
set -eu
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; shift
mkdir -p "$ROOT/build"
[ -f "$ROOT/openrave_b200/csrc/cc_planner.o" ] || g++ -O2 -std=c++20 -fPIC -Wall -c "$ROOT/openrave_b200/csrc/cc_planner.cpp" -o "$ROOT/openrave_b200/csrc/cc_planner.o"
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -shared -Xcompiler -fPIC "$@" \
    -o "$ROOT/build/libcc_$NAME.so" "$ROOT/openrave_b200/csrc/cudacollision.cu" "$ROOT/openrave_b200/csrc/cc_planner.o"
echo "built build/libcc_$NAME.so $*"
