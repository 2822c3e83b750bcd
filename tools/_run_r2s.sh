mkdir -p gpurun_out
for tool in memcheck racecheck initcheck; do
  echo "== compute-sanitizer --tool $tool python tools/sanitize_wavefront.py" >> gpurun_out/r2s_sanitizer.log
  (timeout 1200 compute-sanitizer --tool $tool python tools/sanitize_wavefront.py 2>&1 | tail -6) >> gpurun_out/r2s_sanitizer.log
done
cat gpurun_out/r2s_sanitizer.log
