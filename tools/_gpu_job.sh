mkdir -p gpurun_out
ncu --set full --clock-control none -k regex:k_gemm_f64 -s 12 -c 3 -o gpurun_out/r2_k4_gemm python tools/dense_hv_bench.py 60 300 > /dev/null 2>&1
ncu -i gpurun_out/r2_k4_gemm.ncu-rep --page raw --csv > gpurun_out/r2_k4_gemm_raw.csv 2>/dev/null
python - <<'P'
import csv
rows=list(csv.reader(open('gpurun_out/r2_k4_gemm_raw.csv')))
hdr=rows[0]
keys=[k for k in hdr if any(s in k for s in ['Kernel Name','gpu__time_duration.sum','pipe_tensor','pipe_fp64','dmma','sm__throughput.avg.pct','sm__warps_active.avg.pct','launch__registers','launch__grid_size','smsp__inst_executed_pipe_tensor','l1tex__data_bank_conflicts','smsp__warp_issue_stalled_barrier','smsp__warp_issue_stalled_long_sc','smsp__warp_issue_stalled_short','smsp__warp_issue_stalled_math','smsp__warp_issue_stalled_mio','dram__bytes_read.sum','lts__t_bytes.sum '])]
for r in rows[2:5]:
    for k in keys: print(k, '=', r[hdr.index(k)], rows[1][hdr.index(k)])
    print()
P
