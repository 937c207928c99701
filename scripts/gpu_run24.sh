mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:se_step_lut_global -s 10 -c 2 --csv --log-file gpurun_out/k1c.csv python bench.py --steps 32 --warmup 8 --no-cpu-baseline > gpurun_out/ncu_k1c.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/k1c.csv')))
hdr=None
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); print(d['ID'], d['Kernel Name'][:28], d['Metric Name'], d['Metric Value'], d['Metric Unit'])
PY
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench24.json 2> gpurun_out/bench24.err; tail -3 gpurun_out/bench24.err
python -c "
import json; d=json.load(open('gpurun_out/bench24.json')); print('N=1', d['value'], d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['job_roundtrip']['value'], d['gpu_launches'])"
