mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8) > gpurun_out/r2y_pytest.log
tail -3 gpurun_out/r2y_pytest.log
(timeout 600 python tools/ab_variants.py run 3840x2160x128 2>&1) > gpurun_out/r2y_ab_philox.txt
cat gpurun_out/r2y_ab_philox.txt
for c in 4 5; do (timeout 600 python bench.py --config $c --steps 3 --no-cpu-baseline 2> gpurun_out/r2y_bench_c$c.err | tail -1) > gpurun_out/r2y_bench_c$c.json; cut -c1-160 gpurun_out/r2y_bench_c$c.json; done
rm -f gpurun_out/r2y_sanitizer.log
for tool in memcheck racecheck initcheck; do
  for s in wavefront stream; do
  echo "== compute-sanitizer --tool $tool python tools/sanitize_$s.py" >> gpurun_out/r2y_sanitizer.log
  (timeout 900 compute-sanitizer --tool $tool python tools/sanitize_$s.py 2>&1 | tail -6) >> gpurun_out/r2y_sanitizer.log
  done
done
cat gpurun_out/r2y_sanitizer.log
