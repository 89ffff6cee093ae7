#!/bin/bash
# Build libemplanner.so with different launch-bound / CTA-size defines ON THE GPU BOX and time each.
# usage: tools/tune_variants.sh "<defines 1>" "<defines 2>" ...
set -u
cd "$(dirname "$0")/.."
for defs in "$@"; do
  EMP_DEFINES="$defs" python emplanner_carla_b200/csrc/build.py > /dev/null 2>&1 || { echo "BUILD FAILED: $defs"; continue; }
  for lat in C A; do
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lattice $lat 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%-60s lat=%s value=%9.0f e2e=%9.0f dp=%.3f qp=%.3f ms iters=%s' % ('$defs','$lat',d['value'],d['e2e']['value'],d['kernels_ms']['sl_dp_kernel'],d['kernels_ms']['sl_qp_kernel'],d['config']['qp_iterations']))"
  done
done
