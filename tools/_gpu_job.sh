mkdir -p gpurun_out
timeout 1700 python tools/run_configs.py bqpdual130 2>&1 | grep "^{" | cut -c1-700 | tee gpurun_out/r2_bqpdual130.jsonl
