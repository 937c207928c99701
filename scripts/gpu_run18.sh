mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -s 20 -c 12 --csv --log-file gpurun_out/launches_cfg4.csv python scripts/run_configs.py 4 > gpurun_out/ncu_cfg4.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_cfg4.csv')))
hdr=None
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); print(d['ID'], d['Kernel Name'][:28], d['Metric Name'], d['Metric Value'], d['Metric Unit'])
PY
