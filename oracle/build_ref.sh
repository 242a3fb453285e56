#!/bin/bash
# TEST INFRASTRUCTURE: compiles the *unmodified* reference join path (src/join_base.cpp + headers) from
# where it lies under $GCRE_REF (default /root/reference) into oracle/_ref/ (git-ignored, travels with gpurun).
# Recipe is ours; the reference's own build system (R CMD INSTALL / test/Makefile with clang) is not used.
# Flags follow README.md:27-35 of the reference ("-O3 -march=native") with g++ (clang is not installed), plus
# portable x86-64 level variants because the GPU box's host CPU may differ from this container's.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${GCRE_REF:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -f "$REF/src/join_base.cpp" ]; then
  echo "[build_ref] reference sources not found at $REF - keeping prebuilt files in $OUT" >&2
  exit 0
fi
mkdir -p "$OUT/stub"
: > "$OUT/stub/Rcpp.h"     # join_base.cpp includes <Rcpp.h> but only uses it in a commented-out line
COMMON="-std=c++11 -O3 -mpopcnt -fPIC -shared -pthread -w -I$OUT/stub -I$REF/src"
build() {  # name, arch flags
  g++ $COMMON $2 "$HERE/ref_shim.cpp" "$REF/src/join_base.cpp" -o "$OUT/libgcre_ref_$1.so"
}
build avx512 "-march=x86-64-v4 -mavx512vpopcntdq -mavx512bitalg" &
build avx2   "-march=x86-64-v3" &
build sse42  "-march=x86-64-v2" &
wait
# the reference's own replay harness, unmodified, against the reference join (oracle driver)
g++ -std=c++11 -O3 -march=x86-64-v2 -mpopcnt -pthread -w -I"$OUT/stub" -I"$REF/src" -I"$REF/test" \
    "$REF/test/harness.cpp" "$REF/test/test.cpp" "$REF/src/join_base.cpp" -o "$OUT/ref_harness"
g++ --version | head -1 > "$OUT/COMPILER.txt"
echo "[build_ref] built: $(ls $OUT | tr '\n' ' ')"
# Drop-in proof: the reference's replay harness, UNMODIFIED, compiled against OUR headers (include/gcre/) and linked to
# the CUDA engine instead of the reference's join_base.cpp.  Needs geneticscre_b200/libgcre_b200.so (build it first).
ROOT="$(dirname "$HERE")"
if [ -f "$ROOT/geneticscre_b200/libgcre_b200.so" ]; then
  g++ -std=c++11 -O2 -pthread -w -I"$ROOT/include/gcre" -I"$REF/test" "$REF/test/harness.cpp" "$REF/test/test.cpp" \
      -L"$ROOT/geneticscre_b200" -lgcre_b200 -Wl,-rpath,'$ORIGIN/../../geneticscre_b200' -o "$OUT/b200_harness" \
    && echo "[build_ref] built b200_harness (reference test/harness.cpp on the B200 engine)"
fi
