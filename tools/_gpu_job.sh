# scratch job script for `gpurun -- 'bash tools/_gpu_job.sh'` (edited per experiment); the round's final check was:
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; wc -l gpurun_out/bench_n1.json
python -c "import __graft_entry__ as g; g.smoke()"
