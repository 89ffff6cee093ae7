#!/bin/bash
# ncu --set full capture of one sl_qp_kernel (and one sl_dp_kernel) launch per library variant in tmp_ab/ (or the library
# in place when tmp_ab/ is empty); reports land in gpurun_out/ncu_<name>.ncu-rep, to be read locally with `ncu -i`.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cp emplanner_carla_b200/libemplanner.so /tmp/lib_keep.so
for f in ${NCU_LIBS:-tmp_ab/lib_*.so}; do
  name=$(basename "$f" .so); name=${name#lib_}
  cp "$f" emplanner_carla_b200/libemplanner.so
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:'sl_qp_kernel|sl_dp_kernel' -s 68 -c 2 \
      -f -o gpurun_out/ncu_$name python tools/sched_probe.py --quick > gpurun_out/ncu_$name.log 2>&1
  tail -2 gpurun_out/ncu_$name.log
done
cp /tmp/lib_keep.so emplanner_carla_b200/libemplanner.so
