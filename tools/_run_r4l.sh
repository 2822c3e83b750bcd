mkdir -p gpurun_out
(timeout 1500 python tools/ab_stream.py run 2>&1) > gpurun_out/r4l_ab_stream.txt
cat gpurun_out/r4l_ab_stream.txt
(timeout 900 python -m pytest tests/test_gpu_image.py tests/test_gpu_media.py -m gpu -q -x -k "stream or bvh or media or config4" 2>&1 | tail -3)
