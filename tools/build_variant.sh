#!/usr/bin/env bash
# Build a tuning variant of the library next to the product one:
#   tools/build_variant.sh NAME -DGPUAR_DEC_TOTAL_BIG=1 ...   ->  gpuar_b200/libgpuar_b200_NAME.so
# Select it at run time with GPUAR_B200_LIB=<path> (gpuar_b200/_lib.py).  Variants are scratch:
# they are git-ignored and exist only to A/B a compile-time knob on the GPU box.
set -euo pipefail
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/gpuar_b200/csrc
out=$src/build/variant_$name
mkdir -p "$out"
for f in api host_pipeline encode encode_ws decode index; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC "$@" \
       -c "$src/$f.cu" -o "$out/$f.o" &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$root/gpuar_b200/libgpuar_b200_$name.so" "$out"/*.o
echo "$root/gpuar_b200/libgpuar_b200_$name.so"
