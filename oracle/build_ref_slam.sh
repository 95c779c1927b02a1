#!/bin/sh
# oracle/build_ref_slam.sh -- TEST INFRASTRUCTURE ONLY.  Compiles the reference's own src/slam.cpp and
# src/cone.cpp from where they lie under $REFERENCE into oracle/_ref/ref_slam_replay:
#   * cluon: src/cluon-complete-build.hpp is the single-header library; the reference's CMakeLists.txt:55-69
#     builds its message compiler (cluon-msc) from it and generates opendlv-standard-message-set.{hpp,cpp}
#     from the .odvd -- the same three commands are run here into a scratch directory;
#   * Eigen: the reference's vendored thirdparty/Eigen;
#   * g2o (absent, unpinned upstream): oracle/g2o_facade over the oracle's restated Gauss-Newton
#     (slam_oracle.cpp, built with the reference's Eigen LDLT);
#   * the harness: oracle/ref_slam_replay.cpp.
# Nothing is copied out of the reference tree; outputs go to oracle/_ref/ only (git-ignored).
set -e
REFERENCE=${REFERENCE:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
[ -f "$REFERENCE/src/slam.cpp" ] || { echo "reference tree absent: keeping prebuilt _ref/ (if any)"; exit 0; }
mkdir -p "$OUT"
if [ -x "$OUT/ref_slam_replay" ] && [ "$OUT/ref_slam_replay" -nt "$HERE/ref_slam_replay.cpp" ] && \
   [ "$OUT/ref_slam_replay" -nt "$HERE/g2o_facade/g2o_facade.hpp" ] && [ "$OUT/ref_slam_replay" -nt "$HERE/slam_oracle.cpp" ]; then
  exit 0
fi
echo "building _ref/ref_slam_replay from $REFERENCE/src/slam.cpp"
TMP=$(mktemp -d)
trap 'rm -rf "$TMP"' EXIT
ln -s "$REFERENCE/src/cluon-complete-build.hpp" "$TMP/cluon-complete.hpp"
ln -s "$TMP/cluon-complete.hpp" "$TMP/cluon-complete.cpp"
CXX=${CXX:-g++}
$CXX -o "$TMP/cluon-msc" "$TMP/cluon-complete.cpp" -std=c++14 -pthread -D HAVE_CLUON_MSC -include linux/sockios.h -w
ODVD=$(ls "$REFERENCE"/src/opendlv-standard-message-set-*.odvd | head -1)
"$TMP/cluon-msc" --cpp-sources --cpp-add-include-file=opendlv-standard-message-set.hpp --out="$TMP/opendlv-standard-message-set.cpp" "$ODVD"
"$TMP/cluon-msc" --cpp-headers --out="$TMP/opendlv-standard-message-set.hpp" "$ODVD"
FLAGS="-std=c++14 -O2 -ffp-contract=off -pthread -w -include linux/sockios.h"
INC="-I$TMP -I$HERE/g2o_facade -I$REFERENCE/src -isystem $REFERENCE/thirdparty"
$CXX $FLAGS $INC -c "$REFERENCE/src/slam.cpp" -o "$TMP/slam.o"
$CXX $FLAGS $INC -c "$REFERENCE/src/cone.cpp" -o "$TMP/cone.o"
$CXX $FLAGS $INC -c "$TMP/opendlv-standard-message-set.cpp" -o "$TMP/msgs.o"
$CXX $FLAGS $INC -DORACLE_USE_EIGEN -c "$HERE/slam_oracle.cpp" -o "$TMP/oracle.o"
$CXX $FLAGS $INC -c "$HERE/ref_slam_replay.cpp" -o "$TMP/replay.o"
$CXX -pthread -o "$OUT/ref_slam_replay" "$TMP/replay.o" "$TMP/slam.o" "$TMP/cone.o" "$TMP/msgs.o" "$TMP/oracle.o"
