mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15) > gpurun_out/r2t_pytest.log
tail -6 gpurun_out/r2t_pytest.log
