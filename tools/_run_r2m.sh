mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_image.py -m gpu -q -x -k "f64" 2>&1 | tail -15) > gpurun_out/r2m_pytest.log
(timeout 300 python tools/prof_f64.py 800x600x16 2>&1 | tail -1) > gpurun_out/r2m_f64.jsonl
(timeout 300 python tools/prof_f64.py 3840x2160x8 2>&1 | tail -1) >> gpurun_out/r2m_f64.jsonl
tail -6 gpurun_out/r2m_pytest.log; cat gpurun_out/r2m_f64.jsonl
