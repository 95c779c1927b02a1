#!/bin/sh
# Host code of libslam_b200.so under sanitizers on a box WITHOUT a GPU.  Builds an instrumented copy of the library
# (device code unchanged) into $OUT and preloads tests/stub_cudart.cpp in place of the CUDA runtime (test
# infrastructure: device memory = host heap, kernels = no-ops, nothing computed).
#   host_sanitize.sh asan   AddressSanitizer + UBSan over tests/host_case_stub_runtime.py (default)
#   host_sanitize.sh tsan   ThreadSanitizer over tests/tsan_two_contexts_driver.cpp: two contexts on two host threads
# Any finding aborts with the sanitizer's report; the last line ("{...}" / "ok ...") means clean.
set -e
MODE=${1:-asan}
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
S="$ROOT/opendlv-logic-cfsd18-sensation-slam_b200/csrc"
OUT=${OUT:-/tmp/slam_b200_host_sanitize_$MODE}
mkdir -p "$OUT"
if [ "$MODE" = tsan ]; then
  SAN="-Xcompiler -fsanitize=thread"
else
  SAN="-Xcompiler -fsanitize=address -Xcompiler -fsanitize=undefined -Xcompiler -fno-sanitize-recover=undefined"
fi
for f in capi.cu graph.cu solver.cu assoc.cu symbolic.cpp tileplan.cpp; do
  extra=""; [ "$f" = assoc.cu ] && extra="-fmad=false"
  nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 -Xcompiler -fPIC $SAN $extra -c "$S/$f" -o "$OUT/$f.o" &
done
wait
if [ "$MODE" = tsan ]; then
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/libslam_b200_san.so" "$OUT"/*.o -lcudart -Xcompiler -fsanitize=thread
  g++ -O1 -g -fPIC -shared -fsanitize=thread -o "$OUT/stub_cudart.so" "$ROOT/tests/stub_cudart.cpp"
  g++ -std=c++17 -O1 -g -fsanitize=thread "$ROOT/tests/tsan_two_contexts_driver.cpp" -o "$OUT/two_contexts" \
    -L"$OUT" -l:libslam_b200_san.so -l:stub_cudart.so -Wl,-rpath,"$OUT" -lpthread
  LD_PRELOAD="$OUT/stub_cudart.so" TSAN_OPTIONS=halt_on_error=1 "$OUT/two_contexts"
else
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/libslam_b200_san.so" "$OUT"/*.o -lcudart \
    -Xcompiler -fsanitize=address -Xcompiler -fsanitize=undefined
  g++ -O1 -g -fPIC -shared -fsanitize=address -o "$OUT/stub_cudart.so" "$ROOT/tests/stub_cudart.cpp"
  ASAN_OPTIONS=detect_leaks=0 SLAM_LIB="$OUT/libslam_b200_san.so" SLAM_STUB_CUDART="$OUT/stub_cudart.so" \
    LD_PRELOAD="$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so):$OUT/stub_cudart.so" \
    python "$ROOT/tests/host_case_stub_runtime.py"
fi
