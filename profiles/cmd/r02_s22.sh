# round 2, call 22: v2 dilation, scatter reciprocals once per block, 16 chunks per thread
python -m pytest tests/test_gpu_parity.py -x -q -k "dropblock or dilate or ichan or mc_dropblock_vs" > gpurun_out/r02_s22_pytest.log 2>&1; tail -3 gpurun_out/r02_s22_pytest.log
M=gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $M --clock-control none --profile-from-start off -k regex:dropblock --csv --log-file gpurun_out/r02_s22_mask.csv python tests/prof_step.py 10 2 > gpurun_out/r02_s22.log 2>&1
python - <<'PY'
import csv
rows=list(csv.DictReader([l for l in open('gpurun_out/r02_s22_mask.csv') if not l.startswith('==')]))
d={}
for r in rows:
    d.setdefault((r['ID'], r['Kernel Name'].split('(')[0][:50]),{})[r['Metric Name']]=(r['Metric Value'],r['Metric Unit'])
for k,v in d.items():
    print(k, {m.split('.')[0][:28]:x for m,x in v.items()})
PY
for v in "B2U_DILATE=v1" "B2U_DILATE=v2"; do
  for dt in bf16 fp16; do
  echo "== $v $dt" >> gpurun_out/r02_s22_ab.log
  env $v python bench.py --steps 40 --warmup 5 --dtype $dt --no-e2e --no-cpu --no-train --no-alt 2>> gpurun_out/r02_s22_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
o=d['roofline']['other_kernels_ms_per_step']
print(d['value'], d['ms_per_step'], d['roofline']['sustained_100_steps']['value'], d['clocks']['sm_mhz'], {k:round(v,3) for k,v in o.items() if 'dropblock' in k})" >> gpurun_out/r02_s22_ab.log
  done
done
cat gpurun_out/r02_s22_ab.log; tail -3 gpurun_out/r02_s22_ab.err
