#!/bin/bash
# developer helper: link lib/libdod_<name>.so with csrc/<file>.cu compiled with extra nvcc flags
# usage: tools/build_variant.sh NAME FILE.cu [nvcc flags...]
set -e
cd "$(dirname "$0")/.."
name=$1; src=$2; shift 2
pkg=dinov2-od_b200
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
  -I include "$@" -c $pkg/csrc/$src -o /tmp/variant_$name.o
objs=$(ls $pkg/build/*.o | grep -v "/${src%.cu}.o")
nvcc -shared -o $pkg/lib/libdod_$name.so /tmp/variant_$name.o $objs -cudart static -gencode arch=compute_100a,code=sm_100a
echo $pkg/lib/libdod_$name.so
