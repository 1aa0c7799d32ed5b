mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multiblock.py tests/test_mex_gateway.py -m gpu -q -x 2>&1 | tail -5
