#!/bin/bash
# Builds the drop-in demo (tests/dropin/dropin_demo.cpp) against the reference's headers, once per library it is linked with.
# Only possible where /root/reference exists; the binaries land in tests/dropin/_build/ (git-ignored, they travel to the GPU box).
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
REF=${REF:-/root/reference}
[ -f "$REF/inc/crnlib.h" ] || { echo "no reference headers: keeping prebuilt binaries"; exit 0; }
OUT=$ROOT/tests/dropin/_build
mkdir -p $OUT
FLAGS="-std=c++17 -O1 -include cstdint -I$ROOT/crunch2_b200/csrc/_gen -I$REF/inc $ROOT/tests/dropin/dropin_demo.cpp -lpthread"
[ -f $ROOT/oracle/_ref/liboracle_ref.so ] && g++ $FLAGS -o $OUT/demo_ref -L$ROOT/oracle/_ref -loracle_ref -Wl,-rpath,'$ORIGIN/../../../oracle/_ref'
[ -f $ROOT/tests/cusim/libcrnlib_b200_sim.so ] && g++ $FLAGS -o $OUT/demo_sim -L$ROOT/tests/cusim -lcrnlib_b200_sim -Wl,-rpath,'$ORIGIN/../../cusim'
[ -f $ROOT/crunch2_b200/libcrnlib_b200.so ] && g++ $FLAGS -o $OUT/demo_b200 -L$ROOT/crunch2_b200 -lcrnlib_b200 -Wl,-rpath,'$ORIGIN/../../../crunch2_b200'
exit 0
