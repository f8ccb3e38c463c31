#!/bin/bash
# Developer tool (run under gpurun): A/B the library variants in slam-eds_b200/build on the LM solve.
#   tools/gpu_ab.sh "<sequence counts>" variant...     (variant "product" = slam-eds_b200/libedsgpu.so)
cd "$(dirname "$0")/.."
B=$PWD/slam-eds_b200/build
sizes=$1; shift
for v in "$@"; do
  lib=$B/libedsgpu_$v.so; [ "$v" = product ] && lib=$PWD/slam-eds_b200/libedsgpu.so
  for S in $sizes; do
    EDSGPU_LIBRARY=$lib timeout 40 python tools/lm_time.py $S 10 2>&1 | grep -v "^$" | tail -7 || echo "TIMEOUT $v $S"
  done
done
