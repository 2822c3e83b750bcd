mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -30) > gpurun_out/r2l_pytest.log
(timeout 300 python -m pytest tests/test_render_cli.py -m gpu -q -s 2>&1 | grep -E "denoise\]|passed|failed|Error" | head) > gpurun_out/r2l_misc.log
(timeout 300 python -m rust_pathtracer_b200.render --scene demo --size 800x600 --spp 8 --denoise 4 --out gpurun_out/r2l_demo_den.exr 2>&1 | tail -1) >> gpurun_out/r2l_misc.log
tail -8 gpurun_out/r2l_pytest.log; cat gpurun_out/r2l_misc.log
