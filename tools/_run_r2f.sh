mkdir -p gpurun_out
(timeout 120 python tools/ab_variants.py one 640 360 8 /tmp/none.npy 2>&1 | tail -2) > gpurun_out/r2f_sanity.log
cat gpurun_out/r2f_sanity.log
if grep -q msamples gpurun_out/r2f_sanity.log; then
(timeout 900 python tools/ab_variants.py run 3840x2160x128 2>&1) > gpurun_out/r2f_ab.log
(timeout 900 python tools/ab_variants.py run 1920x1080x256 2>&1) >> gpurun_out/r2f_ab.log
cat gpurun_out/r2f_ab.log
(timeout 900 python -m pytest tests -m gpu -q -x -k "wavefront or benchmarked or tail or image_deterministic or random_scenes or resolved or counters or peer_slot" 2>&1 | tail -15) > gpurun_out/r2f_pytest.log
tail -5 gpurun_out/r2f_pytest.log
fi
