mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_maxcut.py tests/test_mex_gateway.py -m gpu -q -x -k "column or lowdeg or wide_block or gateway" > gpurun_out/r2_pytest_d.log 2>&1; tail -15 gpurun_out/r2_pytest_d.log
timeout 200 python tools/sweep_p.py torus 40,64 > gpurun_out/r2_sweep_torus_gw_b.jsonl 2>&1; cat gpurun_out/r2_sweep_torus_gw_b.jsonl
for g in er torus; do CHK_GRAPH=$g timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/check_colsharded.py 2>&1 | tail -4; done > gpurun_out/r2_colcheck_n2.log 2>&1; cat gpurun_out/r2_colcheck_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; tail -c 3000 gpurun_out/r2_bench_n2.json; tail -5 gpurun_out/r2_bench_n2.err
