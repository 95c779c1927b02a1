#!/bin/sh
# integration/build.sh -- TEST INFRASTRUCTURE.  Proves the drop-in boundary inside the reference tree:
#   1. copies the reference's src/ to a scratch directory and applies integration/slam_b200.patch with `patch -p1`
#      (the committed patch must apply cleanly; integration/make_patched_tree.py regenerates it);
#   2. builds the PATCHED slam.cpp + the reference's cone.cpp + cluon + the generated message set exactly like
#      oracle/build_ref_slam.sh, but linked against libslam_b200.so instead of the g2o facade
#      -> integration/_build/patched_public_replay  (runs on the B200);
#   3. builds the UNMODIFIED reference (g2o answered by oracle/g2o_facade over the oracle) with the same harness
#      -> integration/_build/ref_public_replay       (runs on the CPU, makes the golden file).
# Outputs only under integration/_build/ (git-ignored, travels to the GPU box like the built .so files).
# Nothing is copied out of the reference tree into the repository.
set -e
REFERENCE=${REFERENCE:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(dirname "$HERE")
OUT=$HERE/_build
PKG="$ROOT/opendlv-logic-cfsd18-sensation-slam_b200"
[ -f "$REFERENCE/src/slam.cpp" ] || { echo "reference tree absent: keeping prebuilt integration/_build (if any)"; exit 0; }
mkdir -p "$OUT"
if [ -x "$OUT/patched_public_replay" ] && [ -x "$OUT/ref_public_replay" ] && \
   [ "$OUT/patched_public_replay" -nt "$HERE/slam_b200.patch" ] && [ "$OUT/patched_public_replay" -nt "$HERE/public_api_harness.cpp" ] && \
   [ "$OUT/patched_public_replay" -nt "$ROOT/include/slam_b200.h" ] && [ "$OUT/ref_public_replay" -nt "$HERE/public_api_harness.cpp" ] && \
   [ "$OUT/ref_public_replay" -nt "$ROOT/oracle/slam_oracle.cpp" ]; then
  exit 0
fi
[ -f "$PKG/libslam_b200.so" ] || { echo "libslam_b200.so missing: run __graft_entry__.build() first"; exit 1; }
TMP=$(mktemp -d)
trap 'rm -rf "$TMP"' EXIT
CXX=${CXX:-g++}
# ---- cluon + message set, the three commands of the reference's CMakeLists.txt:55-69 ----
ln -s "$REFERENCE/src/cluon-complete-build.hpp" "$TMP/cluon-complete.hpp"
ln -s "$TMP/cluon-complete.hpp" "$TMP/cluon-complete.cpp"
$CXX -o "$TMP/cluon-msc" "$TMP/cluon-complete.cpp" -std=c++14 -pthread -D HAVE_CLUON_MSC -include linux/sockios.h -w
ODVD=$(ls "$REFERENCE"/src/opendlv-standard-message-set-*.odvd | head -1)
"$TMP/cluon-msc" --cpp-sources --cpp-add-include-file=opendlv-standard-message-set.hpp --out="$TMP/opendlv-standard-message-set.cpp" "$ODVD"
"$TMP/cluon-msc" --cpp-headers --out="$TMP/opendlv-standard-message-set.hpp" "$ODVD"
FLAGS="-std=c++14 -O2 -ffp-contract=off -pthread -w -include linux/sockios.h"
$CXX $FLAGS -I"$TMP" -c "$TMP/opendlv-standard-message-set.cpp" -o "$TMP/msgs.o"
# ---- 1. scratch copy + the committed patch ----
mkdir -p "$TMP/tree/src"
cp "$REFERENCE/src/slam.hpp" "$REFERENCE/src/slam.cpp" "$TMP/tree/src/"
(cd "$TMP/tree" && patch -p1 --no-backup-if-mismatch < "$HERE/slam_b200.patch")
# ---- 2. the patched tree over the C ABI ----
INC="-I$TMP -I$TMP/tree/src -I$REFERENCE/src -I$ROOT/include -isystem $REFERENCE/thirdparty"
$CXX $FLAGS $INC -c "$TMP/tree/src/slam.cpp" -o "$TMP/slam_patched.o"
$CXX $FLAGS $INC -c "$REFERENCE/src/cone.cpp" -o "$TMP/cone.o"
$CXX $FLAGS $INC -c "$HERE/public_api_harness.cpp" -o "$TMP/harness_patched.o"
$CXX -pthread -o "$OUT/patched_public_replay" "$TMP/harness_patched.o" "$TMP/slam_patched.o" "$TMP/cone.o" "$TMP/msgs.o" \
    -L"$PKG" -lslam_b200 -Wl,-rpath,'$ORIGIN/../../opendlv-logic-cfsd18-sensation-slam_b200'
# ---- 3. the unmodified reference over the g2o facade ----
INC="-I$TMP -I$ROOT/oracle/g2o_facade -I$REFERENCE/src -isystem $REFERENCE/thirdparty"
$CXX $FLAGS $INC -c "$REFERENCE/src/slam.cpp" -o "$TMP/slam_ref.o"
$CXX $FLAGS $INC -DORACLE_USE_EIGEN -c "$ROOT/oracle/slam_oracle.cpp" -o "$TMP/oracle.o"
$CXX $FLAGS $INC -c "$HERE/public_api_harness.cpp" -o "$TMP/harness_ref.o"
$CXX -pthread -o "$OUT/ref_public_replay" "$TMP/harness_ref.o" "$TMP/slam_ref.o" "$TMP/cone.o" "$TMP/msgs.o" "$TMP/oracle.o"
echo "built $OUT/patched_public_replay and $OUT/ref_public_replay"
