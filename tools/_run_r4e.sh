mkdir -p gpurun_out
(timeout 1500 python tools/ab_stream.py run 2>&1) > gpurun_out/r4e_ab_stream.txt
cat gpurun_out/r4e_ab_stream.txt
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3)
