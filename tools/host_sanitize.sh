#!/bin/bash
# The HOST code of the library (query builder, chain compiler, certificates' rate bounds, BVH build, mimic-equation compiler, planner
# host) under AddressSanitizer + UndefinedBehaviorSanitizer, driven by the device-free tests.  The device code is not touched by
# this build (compute-sanitizer covers it on the GPU box: tools/gpu_sanitize.py).  usage: tools/host_sanitize.sh [log]
set -eu
ROOT=$(cd "$(dirname "$0")/.." && pwd)
LOG=${1:-/tmp/host_sanitize.log}
mkdir -p "$ROOT/build"
SAN="-fsanitize=address,undefined -fno-omit-frame-pointer"
g++ -O1 -g -std=c++20 -fPIC -Wall $SAN -c "$ROOT/openrave_b200/csrc/cc_planner.cpp" -o "$ROOT/build/cc_planner_asan.o"
/usr/local/cuda/bin/nvcc -O1 -g -std=c++17 -gencode arch=compute_100a,code=sm_100a -shared \
    -Xcompiler -fPIC,-fsanitize=address,-fsanitize=undefined,-fno-omit-frame-pointer \
    -o "$ROOT/build/libcc_asan.so" "$ROOT/openrave_b200/csrc/cudacollision.cu" "$ROOT/build/cc_planner_asan.o" -lasan -lubsan
cd "$ROOT"
TESTS="tests/test_birrt.py tests/test_bvh_build.py tests/test_edge_check.py tests/test_pair_tables.py tests/test_ramp_check.py
       tests/test_cachechecker.py tests/test_affine_dofs.py tests/test_motion_classes.py tests/test_fk_program.py tests/test_mimic_expr.py
       tests/test_broadphase_boxes.py tests/test_capi_symbols.py tests/test_scene_sweep.py"
LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1 \
    CUDACOLLISION_LIB=$ROOT/build/libcc_asan.so python -m pytest $TESTS -q -s -m "not gpu" > "$LOG" 2>&1 || true
tail -2 "$LOG"
echo "sanitizer reports: $(grep -c 'runtime error\|AddressSanitizer' "$LOG" || true)"
rm -f "$ROOT/build/libcc_asan.so" "$ROOT/build/cc_planner_asan.o"      # 30 MB that would otherwise travel to the GPU box
