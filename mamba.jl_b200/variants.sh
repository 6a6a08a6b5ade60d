#!/bin/bash
# Builds kernel variants of one translation unit into mambacuda/variants/lib_<tag>.so (experiments; git-ignored).
# usage: variants.sh <file.cu> <tag> <extra nvcc flags...>
set -e
cd "$(dirname "$0")"
src=$1; tag=$2; shift 2
mkdir -p mambacuda/variants build/var_$tag
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -Xptxas -v -c csrc/$src -o build/var_$tag/${src%.cu}.o 2>&1 | grep -A2 "${KERNEL:-kernel}" | grep "Used\|spill" | head -4
objs=""
for f in build/*.o; do b=$(basename $f); if [ "$b" = "${src%.cu}.o" ]; then objs="$objs build/var_$tag/$b"; else objs="$objs $f"; fi; done
/usr/local/cuda/bin/nvcc -shared -o mambacuda/variants/lib_$tag.so $objs -lcudart
echo built $tag
