mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4) > gpurun_out/r4g_pytest.log; tail -2 gpurun_out/r4g_pytest.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -1)
(timeout 900 python bench.py 2> gpurun_out/r4g_bench.err | tail -1) > gpurun_out/r4g_bench.json; cut -c1-200 gpurun_out/r4g_bench.json
