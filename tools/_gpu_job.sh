mkdir -p gpurun_out
python tools/dual_hv_bench.py 60 64 20
python tools/mb_hv_bench.py 20 20 100 20
timeout 900 python -m pytest tests/test_gpu_affine.py tests/test_gpu_dual.py tests/test_gpu_multiblock.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python tools/run_configs.py bqpdual60 bqp60 2>&1 | cut -c1-260 | tail -2
