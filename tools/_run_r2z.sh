mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x -s 2>&1 | grep -v "^\[build\]" | tail -30) > gpurun_out/r2z_pytest.log
tail -12 gpurun_out/r2z_pytest.log
for c in 4 5; do (timeout 600 python bench.py --config $c --steps 3 --no-cpu-baseline 2> gpurun_out/r2z_bench_c$c.err | tail -1) > gpurun_out/r2z_bench_c$c.json; cut -c1-160 gpurun_out/r2z_bench_c$c.json; done
