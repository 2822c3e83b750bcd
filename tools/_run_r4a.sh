mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6) > gpurun_out/r4a_pytest.log
tail -3 gpurun_out/r4a_pytest.log
rm -f gpurun_out/r4a_configs.jsonl
for c in 4 5; do (timeout 600 python bench.py --config $c --steps 3 --no-cpu-baseline 2> gpurun_out/r4a_bench_c$c.err | tail -1) >> gpurun_out/r4a_configs.jsonl; done
cut -c1-200 gpurun_out/r4a_configs.jsonl
python tools/prof_cfg.py 4 2 3840x2160x2; python tools/prof_cfg.py 4 1 3840x2160x2
