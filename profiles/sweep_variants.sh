#!/bin/bash
# run bench.py once per variant library (copied over the in-tree .so), print the short line
cp scarmapper_b200/libscar_b200.so /tmp/lib_keep.so
for f in "$@"; do
  cp gpurun_variants/$f scarmapper_b200/libscar_b200.so
  touch scarmapper_b200/libscar_b200.so
  timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --kernels-only > gpurun_out/sw_$f.json 2> gpurun_out/sw_$f.err
  python profiles/short.py gpurun_out/sw_$f.json $f
done
cp /tmp/lib_keep.so scarmapper_b200/libscar_b200.so
