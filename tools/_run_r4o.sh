mkdir -p gpurun_out
(timeout 900 python bench.py --no-cpu-baseline 2> gpurun_out/r4o_bench.err | tail -1) > gpurun_out/r4o_bench.json; cut -c1-160 gpurun_out/r4o_bench.json
rm -f gpurun_out/r4o_configs.jsonl
for c in 4 5; do (timeout 600 python bench.py --config $c --steps 3 --no-cpu-baseline 2> gpurun_out/r4o_bench_c$c.err | tail -1) >> gpurun_out/r4o_configs.jsonl; done
cut -c1-140 gpurun_out/r4o_configs.jsonl
