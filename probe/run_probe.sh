#!/bin/bash
# Runs every probe test in its own process (a trap in one must not poison the rest).
cd "$(dirname "$0")"
mkdir -p ../gpurun_out
out=../gpurun_out/probe.log
: > $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $out 2>&1
for t in ${@:-m1 m1s m2 m3 m3s m4 m5 m6 m7 t1 t2 t3 t4 t5 t6 p2 p1}; do
  echo "=== $t" >> $out
  timeout 120 ./umma_probe $t >> $out 2>&1
  echo "exit=$?" >> $out
done
grep -E "RESULT|exit=|===|NOTE|error" $out
