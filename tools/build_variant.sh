#!/bin/bash
# tools/build_variant.sh <name> "<file.cu> [file2.cu ...]" [-D...]: tools/bin/libv_<name>.so = the in-tree objects with
# the listed .cu files recompiled with extra flags
set -e
cd "$(dirname "$0")/.."
name=$1; srcs=$2; shift 2
mkdir -p tools/bin/obj_$name
python kvpress_b200/build.py > /dev/null
for src in $srcs; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
    -c kvpress_b200/csrc/$src -o tools/bin/obj_$name/${src%.cu}.o &
done
wait
objs=""
for o in kvpress_b200/build/*.o; do
  b=$(basename $o)
  if [ -f tools/bin/obj_$name/$b ]; then objs="$objs tools/bin/obj_$name/$b"; else objs="$objs $o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o tools/bin/libv_$name.so $objs
echo tools/bin/libv_$name.so
