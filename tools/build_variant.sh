#!/bin/bash
# tools/build_variant.sh <name> <file.cu> [-D...]: tools/bin/libv_<name>.so = the in-tree objects with one .cu recompiled with extra flags
set -e
cd "$(dirname "$0")/.."
name=$1; src=$2; shift 2
mkdir -p tools/bin/obj_$name
python kvpress_b200/build.py > /dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
  -c kvpress_b200/csrc/$src -o tools/bin/obj_$name/${src%.cu}.o
objs=""
for o in kvpress_b200/build/*.o; do
  b=$(basename $o)
  if [ "$b" = "${src%.cu}.o" ]; then objs="$objs tools/bin/obj_$name/$b"; else objs="$objs $o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o tools/bin/libv_$name.so $objs
echo tools/bin/libv_$name.so
