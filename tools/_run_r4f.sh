mkdir -p gpurun_out
rm -f gpurun_out/r4f_sanitizer.log
for tool in memcheck racecheck initcheck; do
  for s in wavefront stream; do
  echo "== compute-sanitizer --tool $tool python tools/sanitize_$s.py" >> gpurun_out/r4f_sanitizer.log
  (timeout 900 compute-sanitizer --tool $tool python tools/sanitize_$s.py 2>&1 | tail -6) >> gpurun_out/r4f_sanitizer.log
  done
done
cat gpurun_out/r4f_sanitizer.log
for c in 4 5; do (timeout 600 python bench.py --config $c --steps 3 --no-cpu-baseline 2> gpurun_out/r4f_bench_c$c.err | tail -1) > gpurun_out/r4f_bench_c$c.json; cut -c1-160 gpurun_out/r4f_bench_c$c.json; done
