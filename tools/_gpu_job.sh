mkdir -p gpurun_out
MANISDP_K3_WIDE=0 timeout 300 python tools/run_configs.py bqpsparse:20x20 --verbose --opts='{"AL_maxiter": 40}' 2>&1 | grep "^Iter" | cut -c1-150 > gpurun_out/mb_off.txt
MANISDP_K3_WIDE=1 timeout 300 python tools/run_configs.py bqpsparse:20x20 --verbose --opts='{"AL_maxiter": 40}' 2>&1 | grep "^Iter" | cut -c1-150 > gpurun_out/mb_on.txt
paste -d'\n' gpurun_out/mb_off.txt gpurun_out/mb_on.txt | sed 's/, time:.*//' | head -80
