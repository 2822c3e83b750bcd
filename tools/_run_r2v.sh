mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_media.py -q -x 2>&1 | tail -25) > gpurun_out/r2v_media.log
(timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2v_bench.err | tail -1) > gpurun_out/r2v_bench.json
(timeout 1500 python tools/ab_stream.py run 2>&1) > gpurun_out/r2v_ab_stream.txt
tail -12 gpurun_out/r2v_media.log; cut -c1-200 gpurun_out/r2v_bench.json; cat gpurun_out/r2v_ab_stream.txt
