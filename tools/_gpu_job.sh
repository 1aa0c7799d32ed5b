mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multiblock.py -m gpu -q -x 2>&1 | tail -4
python tools/mb_hv_bench.py 20 20 16 20
python tools/mb_hv_bench.py 20 20 4 20
python tools/mb_hv_bench.py 20 20 40 20
timeout 300 python tools/run_configs.py bqpsparse:20x20 2>&1 | cut -c1-300 | tail -1
