#!/bin/bash
# Developer tool (run under gpurun): A/B the library variants in slam-eds_b200/build on the LM solve.
cd "$(dirname "$0")/.."
B=$PWD/slam-eds_b200/build
for v in "$@"; do
  lib=$B/libedsgpu_$v.so; [ "$v" = product ] && lib=$PWD/slam-eds_b200/libedsgpu.so
  for S in 64 8 1; do
    EDSGPU_LIBRARY=$lib timeout 300 python tools/lm_time.py $S 10 2>&1 | grep -v "^$" | tail -6
  done
done
