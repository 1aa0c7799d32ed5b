mkdir -p gpurun_out
python tools/mb_hv_bench.py 20 20 100 20
python tools/mb_hv_bench.py 20 20 16 20
timeout 900 python -m pytest tests/test_gpu_affine.py tests/test_gpu_multiblock.py tests/test_gpu_dual.py tests/test_gpu_baseline_configs.py -m gpu -q -x > gpurun_out/r2_pytest_h.log 2>&1; tail -5 gpurun_out/r2_pytest_h.log
timeout 300 python tools/run_configs.py bqpsparse:20x20 theta112 2>&1 | grep "^{" | cut -c1-330
