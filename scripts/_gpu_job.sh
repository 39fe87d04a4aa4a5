(time python -m pytest tests/test_gpu_weighted.py tests/test_gpu_wide.py tests/test_gpu_pipeline.py tests/test_gpu_tfce.py tests/test_gpu_fullsize.py -x -q) > gpurun_out/r2g_pytest.log 2>&1; tail -4 gpurun_out/r2g_pytest.log
python scripts/probe_wide.py 7 4 300 256 ring1,pipe,pipe_w > gpurun_out/r2g_probe.log 2>&1; tail -3 gpurun_out/r2g_probe.log
TMB_PIPE_ROWS=sell python scripts/probe_wide.py 7 4 300 256 ring1 > gpurun_out/r2g_probe_sell.log 2>&1; tail -1 gpurun_out/r2g_probe_sell.log
