mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -30) > gpurun_out/r2j_pytest.log
(timeout 300 python -m pytest tests/test_gpu_sdf.py -m gpu -q -s 2>&1 | grep -E "sdf image|passed|failed|Error" | head -20) > gpurun_out/r2j_sdf.log
(timeout 300 python -m rust_pathtracer_b200.render --scene sdf --size 960x540 --spp 512 --out gpurun_out/r2j_sdf.png 2>&1 | tail -2) >> gpurun_out/r2j_sdf.log
(timeout 300 python tools/ab_variants.py one 3840 2160 128 /tmp/none.npy 2>&1 | tail -1 | cut -c1-120) >> gpurun_out/r2j_sdf.log
tail -8 gpurun_out/r2j_pytest.log; cat gpurun_out/r2j_sdf.log
