mkdir -p gpurun_out
(timeout 1200 python tools/ab_variants.py run 3840x2160x128 2>&1) > gpurun_out/r2g_ab.log
cat gpurun_out/r2g_ab.log
