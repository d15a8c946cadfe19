#!/bin/bash
# Build the library of an earlier commit next to the in-tree one, for same-box A/B runs (tools/run_gpu_ab.sh loads it with
# RT_LIB=tools/ab_prev_librt.so).  usage: tools/build_prev_lib.sh [<commit>]   (default HEAD: compare against uncommitted work)
set -e
REV=${1:-HEAD}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
cd "$ROOT"
for f in $(git ls-tree -r --name-only "$REV" reptext_b200/csrc include); do
  mkdir -p "$TMP/$(dirname "$f")"; git show "$REV:$f" > "$TMP/$f"
done
cd "$TMP/reptext_b200/csrc"
for f in *.cu; do
  nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
       --expt-relaxed-constexpr -c "$f" -o "${f%.cu}.o" &
done
wait
nvcc -shared -cudart static -gencode arch=compute_100a,code=sm_100a -o "$ROOT/tools/ab_prev_librt.so" *.o -Xlinker --exclude-libs,ALL
rm -rf "$TMP"
echo "built $ROOT/tools/ab_prev_librt.so from $REV"
