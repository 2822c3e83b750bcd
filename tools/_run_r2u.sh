mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_media.py -q -x 2>&1 | tail -25) > gpurun_out/r2u_media.log
(timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15) > gpurun_out/r2u_pytest.log
(timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2u_bench.err | tail -1) > gpurun_out/r2u_bench.json
tail -12 gpurun_out/r2u_media.log; tail -4 gpurun_out/r2u_pytest.log; cut -c1-300 gpurun_out/r2u_bench.json
