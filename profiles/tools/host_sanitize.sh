#!/bin/sh
# Host code of libslam_b200.so under AddressSanitizer + UBSan on a box WITHOUT a GPU: builds an instrumented copy
# of the library (device code unchanged) into /tmp, preloads tests/stub_cudart.cpp in place of the CUDA runtime
# (test infrastructure: device memory = host heap, kernels = no-ops) and runs tests/host_case_stub_runtime.py.
# Any finding aborts with the sanitizer's report; the JSON line at the end means clean.
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
S="$ROOT/opendlv-logic-cfsd18-sensation-slam_b200/csrc"
OUT=${OUT:-/tmp/slam_b200_host_sanitize}
mkdir -p "$OUT"
SAN="-Xcompiler -fsanitize=address -Xcompiler -fsanitize=undefined -Xcompiler -fno-sanitize-recover=undefined"
for f in capi.cu graph.cu solver.cu assoc.cu symbolic.cpp; do
  extra=""; [ "$f" = assoc.cu ] && extra="-fmad=false"
  nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 -Xcompiler -fPIC $SAN $extra -c "$S/$f" -o "$OUT/$f.o" &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/libslam_b200_asan.so" "$OUT"/*.o -lcudart \
  -Xcompiler -fsanitize=address -Xcompiler -fsanitize=undefined
g++ -O1 -g -fPIC -shared -fsanitize=address -o "$OUT/stub_cudart.so" "$ROOT/tests/stub_cudart.cpp"
ASAN_OPTIONS=detect_leaks=0 SLAM_LIB="$OUT/libslam_b200_asan.so" SLAM_STUB_CUDART="$OUT/stub_cudart.so" \
  LD_PRELOAD="$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so):$OUT/stub_cudart.so" \
  python "$ROOT/tests/host_case_stub_runtime.py"
