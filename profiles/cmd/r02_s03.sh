# round 2, call 3: fused conv prologue v2 (metadata prefetch, 12 loads in flight) -- per-layer timing, step A/B
python tests/gpu_diag.py convpro 2>&1 | grep -c "identical True" > gpurun_out/r02_s03_convpro.log
python tests/exp_convpro.py 10 > gpurun_out/r02_s03_exp.log 2>&1
for v in "B2U_FUSED=0" "B2U_FUSED=1" "B2U_FUSE_LEVELS=1,2,3,4"; do
  for dt in bf16; do
    echo "== $v $dt" >> gpurun_out/r02_s03_ab.log
    env $v python bench.py --steps 40 --warmup 5 --dtype $dt --no-e2e --no-cpu --no-train --no-alt 2>> gpurun_out/r02_s03_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['conv_ms_per_step'], d['roofline']['sustained_100_steps']['value'], d['clocks']['sm_mhz'], json.dumps(d['roofline']['other_kernels_ms_per_step']))" >> gpurun_out/r02_s03_ab.log
  done
done
cat gpurun_out/r02_s03_convpro.log; cat gpurun_out/r02_s03_exp.log; cat gpurun_out/r02_s03_ab.log; tail -5 gpurun_out/r02_s03_ab.err
