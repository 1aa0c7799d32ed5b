mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_affine.py tests/test_gpu_edges.py tests/test_sdplib_kats.py tests/test_gpu_baseline_configs.py -m gpu -q -x -k "not n1e6 and not bqp60 and not qs60" > gpurun_out/r2_pytest_f.log 2>&1; tail -4 gpurun_out/r2_pytest_f.log
timeout 200 python tools/theta_hv_bench.py 11 2 8,20,32 50
timeout 200 python tools/theta_hv_bench.py 9 8 20 50
timeout 600 python tools/run_configs.py theta98 theta102 theta112 > gpurun_out/r2_configs_theta.jsonl 2>&1; python - <<'PY'
import json
for l in open('gpurun_out/r2_configs_theta.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); print(d['config'], round(d['seconds'],2), d['iters'], d['hv'], d.get('eig_iters'), d['p_max'], d['obj'], d['eta'], d.get('phase_seconds'))
PY
