#!/bin/bash
# build a variant of the library: only scar_window.cu is recompiled with extra -D flags, the other objects are reused
#   profiles/build_variant.sh NAME "-DFLAG ..."   ->  gpurun_variants/NAME.so
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_variants
NAME=$1; shift
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -cudart static $@ \
     -c scarmapper_b200/csrc/scar_window.cu -o /tmp/scar_window_$NAME.o
OBJS=$(ls scarmapper_b200/csrc/*.o | grep -v scar_window.o)
nvcc -shared -cudart static -gencode arch=compute_100a,code=sm_100a -o gpurun_variants/$NAME.so $OBJS /tmp/scar_window_$NAME.o -lz -lpthread
echo gpurun_variants/$NAME.so
