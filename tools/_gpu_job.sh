mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multiblock.py -m gpu -q -x > gpurun_out/r2_pytest_mb.log 2>&1; tail -25 gpurun_out/r2_pytest_mb.log
timeout 900 python -m pytest tests/test_gpu_affine.py tests/test_gpu_edges.py -m gpu -q -x > gpurun_out/r2_pytest_aff.log 2>&1; tail -5 gpurun_out/r2_pytest_aff.log
