mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "not bqp60" > gpurun_out/r2_pytest_full.log 2>&1; tail -8 gpurun_out/r2_pytest_full.log
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 1500 gpurun_out/r2_bench_n1.json; tail -3 gpurun_out/r2_bench_n1.err
MANISDP_EIG_DEBUG=1 timeout 300 python tools/run_configs.py bqp60 --verbose > gpurun_out/r2_bqp60_profile.log 2>&1; tail -6 gpurun_out/r2_bqp60_profile.log
