mkdir -p gpurun_out
for s in 0 1 2; do timeout 600 python tools/qs60_gpu.py 60 "{\"delta\": 6, \"seed\": $s}" 2>&1 | tail -1 | cut -c1-330; done
