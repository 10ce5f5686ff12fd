# round 2, call 18: per-kernel metrics of the v2 mask pipeline
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,smsp__inst_executed.sum
ncu --metrics $M --clock-control none --profile-from-start off -k regex:dropblock --csv --log-file gpurun_out/r02_s18_mask.csv python tests/prof_step.py 10 2 > gpurun_out/r02_s18.log 2>&1
python - <<'PY'
import csv
rows=list(csv.DictReader([l for l in open('gpurun_out/r02_s18_mask.csv') if not l.startswith('==')]))
d={}
for r in rows:
    d.setdefault((r['ID'], r['Kernel Name'].split('(')[0][:50]),{})[r['Metric Name']]=(r['Metric Value'],r['Metric Unit'])
for k,v in d.items():
    print(k, {m.split('.')[0][:28]:x for m,x in v.items()})
PY
