# round 2, call 7: step-level A/B of the fused prologue v5
for v in "B2U_FUSED=0" "B2U_FUSED=1" "B2U_FUSE_LEVELS=1,2,3,4"; do
  for dt in bf16 fp16; do
    echo "== $v $dt" >> gpurun_out/r02_s07_ab.log
    env $v python bench.py --steps 40 --warmup 5 --dtype $dt --no-e2e --no-cpu --no-train --no-alt 2>> gpurun_out/r02_s07_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['conv_ms_per_step'], d['roofline']['sustained_100_steps']['value'], d['clocks']['sm_mhz'], json.dumps(d['roofline']['other_kernels_ms_per_step']))" >> gpurun_out/r02_s07_ab.log
  done
done
cat gpurun_out/r02_s07_ab.log; tail -5 gpurun_out/r02_s07_ab.err
