mkdir -p gpurun_out
(timeout 120 python tools/ab_variants.py one 640 360 8 /tmp/none.npy 2>&1 | tail -2) > gpurun_out/r2d_sanity.log
cat gpurun_out/r2d_sanity.log
if grep -q msamples gpurun_out/r2d_sanity.log; then
(timeout 900 python tools/ab_variants.py run 3840x2160x128 2>&1) > gpurun_out/r2d_ab.log
cat gpurun_out/r2d_ab.log
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_wavefront -s 1 -c 1 -o gpurun_out/r2d_async python tools/prof_one.py 2 3840x2160x32 > gpurun_out/r2d_ncu.log 2>&1)
ncu -i gpurun_out/r2d_async.ncu-rep --page raw --csv > gpurun_out/r2d_async_raw.csv 2>/dev/null
ncu -i gpurun_out/r2d_async.ncu-rep --page source --csv > gpurun_out/r2d_async_src.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2d_async_raw.csv async2 | grep -E "duration|lanes|issue slots|warp instructions|stall"
fi
