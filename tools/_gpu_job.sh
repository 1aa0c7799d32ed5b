mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_maxcut.py -m gpu -q -k "lowdeg or wide_block" > gpurun_out/r2_pytest_c.log 2>&1; tail -8 gpurun_out/r2_pytest_c.log
timeout 200 python tools/sweep_p.py torus 40,64 > gpurun_out/r2_sweep_torus_gw.jsonl 2>&1; cat gpurun_out/r2_sweep_torus_gw.jsonl
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_spmm_lowdeg -s 30 -c 1 -f -o gpurun_out/r2_lowdeg_torus_p64 python tools/sweep_p.py torus 64 > gpurun_out/ncu_lowdeg.log 2>&1; tail -2 gpurun_out/ncu_lowdeg.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_bm_ -s 100 -c 5 -f -o gpurun_out/r2_bm_er_p64 python tools/sweep_p.py er 64 > gpurun_out/ncu_bm.log 2>&1; tail -2 gpurun_out/ncu_bm.log
timeout 900 python tools/run_configs.py er:1000000 --verbose --opts='{"p0": 256, "delta": 24}' > gpurun_out/r2_c5_er1e6_p256.log 2>&1; tail -12 gpurun_out/r2_c5_er1e6_p256.log
