#!/bin/bash
# A/B of prebuilt library variants for the S-T kernels (tmp_ab/lib_*.so)
cd "$(dirname "$0")/.."
cp emplanner_carla_b200/libemplanner.so /tmp/lib_keep.so
for f in tmp_ab/lib_*.so; do
  cp "$f" emplanner_carla_b200/libemplanner.so
  echo "== $f"; python tools/st_bench.py 4096; python tools/st_back_bench.py 4096 3 2>&1 | sed -n 1p
done
cp /tmp/lib_keep.so emplanner_carla_b200/libemplanner.so
