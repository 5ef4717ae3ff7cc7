#!/usr/bin/env bash
# oracle/build_ref.sh — TEST INFRASTRUCTURE.  Compiles the reference planner sources *where they
# lie* under $REF (default /root/reference) into oracle/_ref/:
#   libclrrt_ref.so          unmodified reference (+ ROS shim, + collision hook, see ref_driver.cpp)
#   libclrrt_ref_defined.so  same, from sed-patched TEMPORARY copies of three files that remove the
#                            reference's undefined behaviour on the hot path:
#     controller.cpp:39, simulation.cpp:66   ref.v[IDwp+LAlong]      -> index clamped to size()-1
#     controller.cpp:76                      IDwp==ref.x.size()      -> IDwp==ref.x.size()-1
#     simulation.cpp:28                      Euler loop i<=x.size()  -> i<7 (dx has 7 entries)
#     old_collisioncheck.cpp:75              normsY[3] never written -> normsY[3]=0
# Flags: -O3 -DNDEBUG (= CMake Release with GCC, which the reference's README.md:59-64 recommends),
# no -march=native, no fast-math.  Nothing is copied into the repository; oracle/_ref/ is git-ignored.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/rrt/src" ]; then
  echo "build_ref: $REF/rrt/src not present (GPU box?) — keeping prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
CXX="${CXX:-g++}"
FLAGS="-std=c++11 -O3 -DNDEBUG -fPIC -shared -w -I$HERE/shim -I$REF/rrt/include"
$CXX $FLAGS -I"$REF/rrt/src" "$HERE/ref_driver.cpp" -o "$OUT/libclrrt_ref.so"

TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
sed -e 's/ref\.v\[IDwp+LAlong\]/ref.v[std::min<int>(IDwp+LAlong,(int)ref.v.size()-1)]/' \
    -e 's/IDwp==ref\.x\.size()/IDwp==ref.x.size()-1/' \
    "$REF/rrt/src/controller.cpp" > "$TMP/controller.cpp"
sed -e 's/ref\.v\[control\.IDwp+LAlong\]/ref.v[std::min<int>(control.IDwp+LAlong,(int)ref.v.size()-1)]/' \
    -e 's/i<= x\.size()/i<7/' \
    "$REF/rrt/src/simulation.cpp" > "$TMP/simulation.cpp"
sed -e 's/normsX\[3\] = -(verticesX\[0\]-verticesX\[3\]);/normsX[3] = -(verticesX[0]-verticesX[3]); normsY[3] = 0;/' \
    "$REF/rrt/src/old_collisioncheck.cpp" > "$TMP/old_collisioncheck.cpp"
# every patch must have applied exactly where intended
grep -q 'std::min<int>(IDwp+LAlong' "$TMP/controller.cpp"
grep -q 'IDwp==ref.x.size()-1' "$TMP/controller.cpp"
grep -q 'std::min<int>(control.IDwp+LAlong' "$TMP/simulation.cpp"
grep -q 'i<7' "$TMP/simulation.cpp"
grep -q 'normsY\[3\] = 0' "$TMP/old_collisioncheck.cpp"
$CXX $FLAGS -DREF_DEFINED -I"$TMP" -I"$REF/rrt/src" "$HERE/ref_driver.cpp" -o "$OUT/libclrrt_ref_defined.so"
echo "build_ref: built $OUT/libclrrt_ref.so and libclrrt_ref_defined.so"
