mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt 2>&1
nproc >> gpurun_out/r2a_gpu.txt
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60) > gpurun_out/r2a_pytest.log
(timeout 600 python bench.py 2> gpurun_out/r2a_bench.err | tail -1) > gpurun_out/r2a_bench.json
(timeout 600 python tools/function_parity.py run 2>&1 | tail -120) > gpurun_out/r2a_fnparity.log
(timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2a_ncu_bench.log 2>&1)
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_wavefront -s 1 -c 1 -o gpurun_out/r2a_wavefront python tools/prof_one.py 2 3840x2160x32 > gpurun_out/r2a_ncu_full.log 2>&1)
ncu -i gpurun_out/r2a_wavefront.ncu-rep --page raw --csv > gpurun_out/r2a_wavefront_raw.csv 2>/dev/null
grep -E "passed|failed" gpurun_out/r2a_pytest.log | tail -3; cat gpurun_out/r2a_bench.json; tail -3 gpurun_out/r2a_bench.err
