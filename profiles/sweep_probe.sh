#!/bin/bash
# profiles/sweep_probe.sh "probe args" VARIANT.so ...  : profiles/probe.py once per variant library
ARGS=$1; shift
cp scarmapper_b200/libscar_b200.so /tmp/lib_keep.so
for f in "$@"; do
  cp gpurun_variants/$f scarmapper_b200/libscar_b200.so
  echo -n "$f "; timeout 300 python profiles/probe.py $ARGS 2>&1 | tail -1 | cut -c1-110
done
cp /tmp/lib_keep.so scarmapper_b200/libscar_b200.so
