for f in BASE NOTAPS NOACC EXPERIMENT_FP32_GEOM; do
cp slam-eds_b200/libedsgpu_t_$f.so slam-eds_b200/libedsgpu_timing.so
echo "== $f"; python scratch_timing.py gen3_vga 2>&1 | tail -1 | cut -c1-140; python scratch_timing_batch.py 64 2>&1 | tail -1 | cut -c1-140
done
