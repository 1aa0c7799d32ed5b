mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dual.py -m gpu -q -x > gpurun_out/r2_pytest_dual.log 2>&1; tail -40 gpurun_out/r2_pytest_dual.log
timeout 900 python -m pytest tests/test_gpu_multiblock.py tests/test_mex_gateway.py tests/test_gpu_affine.py -m gpu -q -x > gpurun_out/r2_pytest_mb.log 2>&1; tail -5 gpurun_out/r2_pytest_mb.log
