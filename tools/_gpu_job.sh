mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_full5.log 2>&1; tail -5 gpurun_out/r2_pytest_full5.log
timeout 900 python bench.py > gpurun_out/r2_bench_n1_d.json 2> gpurun_out/r2_bench_n1_d.err; wc -l gpurun_out/r2_bench_n1_d.json; tail -1 gpurun_out/r2_bench_n1_d.err | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()"
