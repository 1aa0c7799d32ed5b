mkdir -p gpurun_out
timeout 600 python tools/run_configs.py bqp60 g1 bqp20 theta98 theta102 > gpurun_out/r2_configs_adaptive.jsonl 2>&1; cut -c1-900 gpurun_out/r2_configs_adaptive.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 60 --csv --log-file gpurun_out/r2_theta_launches.csv python tools/theta_hv_bench.py 11 2 20 20 > /dev/null 2>&1; python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_theta_launches.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
from collections import OrderedDict
agg=OrderedDict()
for r in rows[1:]:
    k=r[ik][:70]; agg.setdefault(k,[]).append(float(r[iv].replace(',','')))
for k,v in agg.items(): print(len(v), round(sum(v)/len(v)/1000,2),'us', k)
PY
timeout 900 python -m pytest tests -m gpu -q -x -k "full_solve or config4 or config3 or kkt or known_answer or sdplib or sharded or column" > gpurun_out/r2_pytest_e.log 2>&1; tail -5 gpurun_out/r2_pytest_e.log
