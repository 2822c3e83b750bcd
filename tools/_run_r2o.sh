mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12) > gpurun_out/r2o_pytest.log
(timeout 300 python tools/prof_f64.py 3840x2160x16 2>&1 | tail -1) > gpurun_out/r2o_f64.json
(timeout 300 python tools/prof_f64.py 800x600x64 2>&1 | tail -1) >> gpurun_out/r2o_f64.json
tail -4 gpurun_out/r2o_pytest.log; cat gpurun_out/r2o_f64.json
