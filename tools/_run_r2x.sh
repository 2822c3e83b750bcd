mkdir -p gpurun_out
(timeout 1500 python tools/ab_stream.py run 2>&1) > gpurun_out/r2x_ab_stream.txt
cat gpurun_out/r2x_ab_stream.txt
