mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40) > gpurun_out/r2_pytest2.log
(timeout 900 python tools/function_parity.py run 2>&1 | tail -80) > gpurun_out/r2_fnparity2.log
(timeout 600 python tools/ab_variants.py run 3840x2160x128 2>&1) > gpurun_out/r2_ab2.log
(PTB200_LIB=$PWD/rust_pathtracer_b200/libptb200_strict.so timeout 300 python tools/ab_variants.py one 3840 2160 32 /tmp/none.npy 2>&1 | tail -1) > gpurun_out/r2_strict_perf.log
(timeout 600 python tools/parity_outliers.py 1920 1080 2 2>&1 | tail -3) > gpurun_out/r2_outliers.jsonl
(timeout 600 python tools/parity_outliers.py 3840 2160 1 2>&1 | tail -3) >> gpurun_out/r2_outliers.jsonl
tail -5 gpurun_out/r2_pytest2.log; cat gpurun_out/r2_ab2.log gpurun_out/r2_strict_perf.log gpurun_out/r2_outliers.jsonl
