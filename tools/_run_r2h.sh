mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/r2h_pytest.log
(timeout 600 python bench.py 2> gpurun_out/r2h_bench.err | tail -1) > gpurun_out/r2h_bench.json
for c in 1 2 4 5; do (timeout 600 python bench.py --config $c --steps 3 2> gpurun_out/r2h_bench_c$c.err | tail -1) > gpurun_out/r2h_bench_c$c.json; done
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_wavefront -s 1 -c 1 -o gpurun_out/r2h_wavefront python tools/prof_one.py 2 3840x2160x32 > gpurun_out/r2h_ncu_full.log 2>&1)
ncu -i gpurun_out/r2h_wavefront.ncu-rep --page raw --csv > gpurun_out/r2h_wavefront_raw.csv 2>/dev/null
ncu -i gpurun_out/r2h_wavefront.ncu-rep --page source --csv > gpurun_out/r2h_wavefront_src.csv 2>/dev/null
grep -E "passed|failed" gpurun_out/r2h_pytest.log | tail -3; cat gpurun_out/r2h_bench.json | cut -c1-300; tail -3 gpurun_out/r2h_bench.err
for c in 1 2 4 5; do cut -c1-200 gpurun_out/r2h_bench_c$c.json; tail -2 gpurun_out/r2h_bench_c$c.err; done
