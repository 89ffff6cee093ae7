#!/bin/bash
# profiles/sass_r2.txt: per-cubin opcode counts that show what the library is made of (TMA bulk copies, mbarrier waits,
# FP64 math, shuffles, system-scope release/acquire; no tensor-core ops - the path is banded FP64).
cd "$(dirname "$0")/.."
tmp=$(mktemp -d)
( cd $tmp && cuobjdump -xelf all "$OLDPWD/emplanner_carla_b200/libemplanner.so" > /dev/null )
echo "libemplanner.so  $(sha256sum emplanner_carla_b200/libemplanner.so | cut -c1-16)  built from $(git rev-parse --short HEAD) (+ working tree)"
cuobjdump -lelf emplanner_carla_b200/libemplanner.so | sed 's/^/  /'
for c in $tmp/*.cubin; do
  s=$(cuobjdump -sass $c)
  printf "%-28s" "$(basename $c)"
  for op in UBLKCP SYNCS DFMA DMUL DADD DSETP MUFU SHFL LDS STS ATOMG MEMBAR "STG.*STRONG.SYS" "LDG.*STRONG.SYS" "MEMBAR.*SYS" HMMA UTCHMMA QMMA; do
    printf " %s=%s" "$op" "$(echo "$s" | grep -cE "^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9] )?$op")"
  done
  echo
done
rm -rf $tmp
