mkdir -p gpurun_out
MANISDP_EIG_DEBUG=1 timeout 300 python tools/run_configs.py theta112 2>&1 | grep -v "^{" | tail -3
MANISDP_EIG_DEBUG=1 timeout 300 python tools/run_configs.py bqp60 2>&1 | tail -3 | cut -c1-400
