#!/bin/bash
# A/B of prebuilt library variants (tmp_ab/lib_<name>.so, built locally): serial / pipelined / e2e throughput of each.
cd "$(dirname "$0")/.."
cp emplanner_carla_b200/libemplanner.so /tmp/lib_keep.so
for f in tmp_ab/lib_*.so; do
  cp "$f" emplanner_carla_b200/libemplanner.so
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%-28s value=%9.0f serial=%9.0f e2e=%9.0f dp=%.3f qp=%.3f ms' % ('$f',d['value'],d['pipeline']['value_serial_steps'],d['e2e']['value'],d['kernels_ms']['sl_dp_kernel'],d['kernels_ms']['sl_qp_kernel']))"
done
cp /tmp/lib_keep.so emplanner_carla_b200/libemplanner.so
