mkdir -p gpurun_out
for v in "" f64_t320_p960 f64_t384_p1152 f64_t192_p1152 f64_t512_p1024; do
  if [ -n "$v" ]; then export PTB200_LIB=$PWD/rust_pathtracer_b200/variants/libptb200_$v.so; fi
  echo "$v $(timeout 300 python tools/prof_f64.py 3840x2160x8 2>&1 | tail -1)" >> gpurun_out/r2n_f64.txt
done
cat gpurun_out/r2n_f64.txt
