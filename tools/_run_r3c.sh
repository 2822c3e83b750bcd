mkdir -p gpurun_out
# full captures of the bounce-1 shade, finish and trace kernels of config 4 (launch order per bounce >= 1: trace, finish, shade, trace<ANY>)
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_stream_shade|k_stream_finish" -s 2 -c 3 -o gpurun_out/r3c_shade python tools/prof_cfg.py 4 3 3840x2160x1 > gpurun_out/r3c_ncu.log 2>&1)
ncu -i gpurun_out/r3c_shade.ncu-rep --page raw --csv > gpurun_out/r3c_shade_raw.csv 2>/dev/null
ncu -i gpurun_out/r3c_shade.ncu-rep --page source --csv > gpurun_out/r3c_shade_src.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r3c_shade_raw.csv "config 4 shade / finish" | head -80
