mkdir -p gpurun_out
for v in "" trace_mb6 trace_mb8; do
  if [ -n "$v" ]; then export PTB200_LIB=$PWD/rust_pathtracer_b200/variants/libptb200_$v.so; fi
  echo "$v cfg4 $(timeout 300 python tools/prof_cfg.py 4 3 3840x2160x4 2>&1 | tail -1) cfg5 $(timeout 300 python tools/prof_cfg.py 5 3 3840x2160x8 2>&1 | tail -1)" >> gpurun_out/r2q.txt
done
cat gpurun_out/r2q.txt
