mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_baseline_configs.py tests/test_gpu_multiblock.py -m gpu -q -x 2>&1 | tail -15
