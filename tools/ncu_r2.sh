#!/bin/bash
# Developer tool (run under gpurun): the round-2 `ncu --set full` captures behind profiles/r02_*.md.
cd "$(dirname "$0")/.."
N="ncu --set full --clock-control none --import-source on -f"
$N -k regex:track_lm_kernel --launch-skip 2 --launch-count 1 -o gpurun_out/prof_r2_lm python tools/lm_time.py 64 2 > gpurun_out/ncu_r2_lm.log 2>&1
$N -k regex:scatter_events_kernel --launch-count 1 -o gpurun_out/prof_r2_scatter python tools/lm_time.py 64 1 > gpurun_out/ncu_r2_ef.log 2>&1
$N -k regex:blur_norm_kernel --launch-count 1 -o gpurun_out/prof_r2_blur python tools/lm_time.py 64 1 >> gpurun_out/ncu_r2_ef.log 2>&1
$N -k regex:ba_lin_top_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/prof_r2_ba_lin_top python tools/ba_bench.py > gpurun_out/ncu_r2_ba.log 2>&1
$N -k regex:ba_top_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/prof_r2_ba_top python tools/ba_bench.py >> gpurun_out/ncu_r2_ba.log 2>&1
$N -k regex:coarse_track_kernel --launch-skip 1 --launch-count 1 -o gpurun_out/prof_r2_coarse_track python tools/coarse_bench.py > gpurun_out/ncu_r2_coarse.log 2>&1
ls -la gpurun_out/prof_r2_*.ncu-rep
