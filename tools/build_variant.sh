#!/bin/bash
# Developer tool: build an experimental variant of the library next to the product build.
#   tools/build_variant.sh <name> [extra nvcc flags...]   ->  slam-eds_b200/build/libedsgpu_<name>.so
# Select it at run time with EDSGPU_LIBRARY=<path>.  SRC_DIR=<dir> builds another source tree's csrc/ (e.g. a git worktree).
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=${SRC_DIR:-$root/slam-eds_b200}
mkdir -p "$root/slam-eds_b200/build"
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall,-Wno-unused-function \
  --expt-relaxed-constexpr "$@" -I"$root/include" -shared -o "$root/slam-eds_b200/build/libedsgpu_$name.so" \
  "$src"/csrc/context.cu "$src"/csrc/event_frame.cu "$src"/csrc/tracker.cu "$src"/csrc/ba.cu "$src"/csrc/coarse.cu "$src"/csrc/depth.cu -lcudart
echo "built slam-eds_b200/build/libedsgpu_$name.so"
