mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_affine.py tests/test_gpu_dual.py -m gpu -q -x -k "beyond_512" 2>&1 | tail -12
