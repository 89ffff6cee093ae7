#!/bin/bash
# Build libemplanner.so with different defines ON THE GPU BOX and report serial/pipelined throughput + DP instruction count.
set -u
cd "$(dirname "$0")/.."
for defs in "$@"; do
  EMP_DEFINES="$defs" python emplanner_carla_b200/csrc/build.py > /dev/null 2>&1 || { echo "BUILD FAILED: $defs"; continue; }
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%-40s value=%9.0f serial=%9.0f e2e=%9.0f dp=%.3f qp=%.3f ms' % ('$defs',d['value'],d['pipeline']['value_serial_steps'],d['e2e']['value'],d['kernels_ms']['sl_dp_kernel'],d['kernels_ms']['sl_qp_kernel']))"
done
