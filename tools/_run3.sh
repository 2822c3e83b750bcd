mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -150) > gpurun_out/r2_pytest3.log
(timeout 600 python tools/ab_variants.py run 3840x2160x128 2>&1) > gpurun_out/r2_ab3.log
(timeout 600 python bench.py --steps 4 2> gpurun_out/r2_bench3.err | tail -1) > gpurun_out/r2_bench3.json
(timeout 300 python bench.py --config 1 --steps 2 2> gpurun_out/r2_bench3_c1.err | tail -1) > gpurun_out/r2_bench3_c1.json
(timeout 300 python bench.py --config 2 --steps 2 2> gpurun_out/r2_bench3_c2.err | tail -1) > gpurun_out/r2_bench3_c2.json
grep -E "passed|failed" gpurun_out/r2_pytest3.log | tail -3; cat gpurun_out/r2_ab3.log; cat gpurun_out/r2_bench3.json; tail -3 gpurun_out/r2_bench3.err; cat gpurun_out/r2_bench3_c1.json; tail -3 gpurun_out/r2_bench3_c1.err; cat gpurun_out/r2_bench3_c2.json
