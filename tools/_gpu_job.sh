mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_full3.log 2>&1; tail -6 gpurun_out/r2_pytest_full3.log
timeout 900 python bench.py > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_bench_n1_b.err; tail -c 3000 gpurun_out/r2_bench_n1_b.json; tail -3 gpurun_out/r2_bench_n1_b.err
