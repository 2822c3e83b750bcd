mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt 2>&1
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/r2_pytest1.log
(timeout 600 python tools/function_parity.py run 2>&1 | tail -80) > gpurun_out/r2_fnparity.log
(timeout 300 python tools/ab_variants.py run 3840x2160x128 2>&1) > gpurun_out/r2_ab1.log
for v in "" rust_pathtracer_b200/variants/libptb200_r1_nan.so; do
  if [ -n "$v" ]; then export PTB200_LIB=$PWD/$v; fi
  timeout 300 python tools/nan_rate.py 3840 2160 1024 3 >> gpurun_out/r2_nan.jsonl 2>&1
done
unset PTB200_LIB
tail -3 gpurun_out/r2_pytest1.log; cat gpurun_out/r2_ab1.log; cat gpurun_out/r2_nan.jsonl
