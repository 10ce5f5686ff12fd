# round 2, call 32: which levels should fuse, now that the plain MT = 2 kernels are faster?  (same box)
for rep in 1 2; do
for v in "B2U_FUSE_LEVELS=0,1,2,3,4" "B2U_FUSE_LEVELS=1,2,3,4" "B2U_FUSE_LEVELS=2,3,4" "B2U_FUSED=0"; do
  for dt in fp16; do
  echo "== $v $dt" >> gpurun_out/r02_s32_ab.log
  env $v python bench.py --steps 40 --warmup 5 --dtype $dt --no-e2e --no-cpu --no-train --no-alt 2>> gpurun_out/r02_s32_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['conv_ms_per_step'], d['roofline']['sustained_100_steps']['value'], d['clocks']['sm_mhz'])" >> gpurun_out/r02_s32_ab.log
  done
done
done
cat gpurun_out/r02_s32_ab.log; tail -3 gpurun_out/r02_s32_ab.err
