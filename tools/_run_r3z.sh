mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/r3z_pytest.log
tail -3 gpurun_out/r3z_pytest.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -3) > gpurun_out/r3z_smoke.log; tail -1 gpurun_out/r3z_smoke.log
(timeout 900 python bench.py 2> gpurun_out/r3z_bench.err | tail -1) > gpurun_out/r3z_bench.json; cut -c1-250 gpurun_out/r3z_bench.json
(timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/r3z_ref.err | tail -1) > gpurun_out/r3z_ref.json; cut -c1-250 gpurun_out/r3z_ref.json
rm -f gpurun_out/r3z_configs.jsonl
for c in 1 2 4 5; do (timeout 600 python bench.py --config $c --steps 3 --no-cpu-baseline 2> gpurun_out/r3z_bench_c$c.err | tail -1) >> gpurun_out/r3z_configs.jsonl; done
cut -c1-200 gpurun_out/r3z_configs.jsonl
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3z_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r3z_ncu_bench.log 2>&1); grep -c k_render gpurun_out/r3z_launches.csv
python -m rust_pathtracer_b200.render --scene media --size 960x540 --spp 1024 --out gpurun_out/media_960x540_1024spp.png 2>&1 | tail -1
python -m rust_pathtracer_b200.render --scene lights --size 960x540 --spp 1024 --out gpurun_out/lights_960x540_1024spp.png 2>&1 | tail -1
