#!/bin/bash
# A/B of prebuilt library variants (tmp_ab/lib_<name>.so, built locally with different defines): kernel times of each
# on the benchmark batch (tools/sched_probe.py --quick).  The library in place is restored afterwards.
cd "$(dirname "$0")/.."
cp emplanner_carla_b200/libemplanner.so /tmp/lib_keep.so
for f in tmp_ab/lib_*.so; do
  echo "== $f"
  cp "$f" emplanner_carla_b200/libemplanner.so
  python tools/sched_probe.py --quick 2>&1 | tail -3
done
cp /tmp/lib_keep.so emplanner_carla_b200/libemplanner.so
