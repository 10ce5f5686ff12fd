# round 2, call 23: do shorter mask blocks help the forward start sooner?  centers trips per block 0 (whole) / 12 / 18
for v in "B2U_CENTERS_SHORT=0" "B2U_CENTERS_SHORT=1 B2U_CENTERS_TPB=12" "B2U_CENTERS_SHORT=1 B2U_CENTERS_TPB=18"; do
  for dt in bf16 fp16; do
  echo "== $v $dt" >> gpurun_out/r02_s23_ab.log
  env $v python bench.py --steps 40 --warmup 5 --dtype $dt --no-e2e --no-cpu --no-train --no-alt 2>> gpurun_out/r02_s23_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
o=d['roofline']['other_kernels_ms_per_step']
print(d['value'], d['ms_per_step'], d['roofline']['sustained_100_steps']['value'], d['clocks']['sm_mhz'], {k:round(v,3) for k,v in o.items() if 'dropblock' in k})" >> gpurun_out/r02_s23_ab.log
  done
done
cat gpurun_out/r02_s23_ab.log; tail -3 gpurun_out/r02_s23_ab.err
