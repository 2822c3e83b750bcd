mkdir -p gpurun_out
(timeout 900 python bench.py 2> gpurun_out/r4j_bench.err | tail -1) > gpurun_out/r4j_bench.json; cut -c1-200 gpurun_out/r4j_bench.json
rm -f gpurun_out/r4j_configs.jsonl
for c in 1 2; do (timeout 600 python bench.py --config $c --steps 3 --no-cpu-baseline 2> gpurun_out/r4j_bench_c$c.err | tail -1) >> gpurun_out/r4j_configs.jsonl; done
cut -c1-160 gpurun_out/r4j_configs.jsonl
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r4j_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r4j_ncu_bench.log 2>&1); grep -c k_render gpurun_out/r4j_launches.csv
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_wavefront -s 1 -c 1 -o gpurun_out/r4j_wavefront python tools/prof_one.py 2 3840x2160x32 > gpurun_out/r4j_ncu_full.log 2>&1)
ncu -i gpurun_out/r4j_wavefront.ncu-rep --page raw --csv > gpurun_out/r4j_wavefront_raw.csv 2>/dev/null
ncu -i gpurun_out/r4j_wavefront.ncu-rep --page source --csv > gpurun_out/r4j_wavefront_src.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r4j_wavefront_raw.csv "r02-e" | grep -E "launch 0|duration|lanes|issue|stall|registers|DRAM|warp instructions"
