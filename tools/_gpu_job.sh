mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_mex_gateway.py -m gpu -q -x > gpurun_out/r2_pytest_g.log 2>&1; tail -15 gpurun_out/r2_pytest_g.log
