mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_affine.py tests/test_gpu_maxcut.py tests/test_gpu_edges.py tests/test_gpu_dual.py -m gpu -q -x 2>&1 | tail -3
MANISDP_EIG_DEBUG=1 timeout 600 python tools/qs60_gpu.py 60 '{"delta": 6, "seed": 2}' 2>&1 | grep -v "manisdp rank" | tail -2 | cut -c1-300
MANISDP_EIG_DEBUG=1 timeout 300 python tools/run_configs.py theta112 theta102 bqp60 bqpdual60 2>&1 | grep -v "manisdp rank" | cut -c1-230 | tail -8
