mkdir -p gpurun_out
(timeout 1500 python tools/converged_parity.py 3840 2160 8 128 2>&1 | tail -1) > gpurun_out/r2k_converged_4k_1024.json
(timeout 600 python tools/converged_parity.py 960 540 8 512 2>&1 | tail -1) > gpurun_out/r2k_converged_960_4096.json
(timeout 600 python tools/nan_rate.py 3840 2160 1024 3 2>&1 | tail -1) > gpurun_out/r2k_nan.json
(timeout 300 python -m pytest tests/test_gpu_image.py -m gpu -q -s -k "config4 or benchmarked" 2>&1 | grep -E "parity\]|passed|failed") > gpurun_out/r2k_parity_prints.log
cat gpurun_out/r2k_converged_4k_1024.json gpurun_out/r2k_converged_960_4096.json gpurun_out/r2k_nan.json gpurun_out/r2k_parity_prints.log
