mkdir -p gpurun_out
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_wavefront -s 1 -c 1 -o gpurun_out/r2c_async python tools/prof_one.py 2 3840x2160x32 > gpurun_out/r2c_ncu.log 2>&1)
ncu -i gpurun_out/r2c_async.ncu-rep --page raw --csv > gpurun_out/r2c_async_raw.csv 2>/dev/null
ncu -i gpurun_out/r2c_async.ncu-rep --page source --csv > gpurun_out/r2c_async_src.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2c_async_raw.csv async
