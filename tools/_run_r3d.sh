mkdir -p gpurun_out
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_stream_trace" -s 2 -c 2 -o gpurun_out/r3d_trace python tools/prof_cfg.py 4 3 3840x2160x1 > gpurun_out/r3d_ncu.log 2>&1)
ncu -i gpurun_out/r3d_trace.ncu-rep --page raw --csv > gpurun_out/r3d_trace_raw.csv 2>/dev/null
ncu -i gpurun_out/r3d_trace.ncu-rep --page source --csv > gpurun_out/r3d_trace_src.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r3d_trace_raw.csv "config 4 trace" | grep -E "launch|duration|lanes|issue|stall|registers|occupancy"
