# round 2, call 26: what does the mask build cost the captured step?  (timing only: masks are stale after the first pair)
for v in "B2U_EXP_SKIP_MASKS=" "B2U_EXP_SKIP_MASKS=all" "B2U_EXP_SKIP_MASKS=centers" "B2U_EXP_SKIP_MASKS=dilate"; do
  for dt in fp16; do
  echo "== $v $dt" >> gpurun_out/r02_s26_ab.log
  env $v python bench.py --steps 40 --warmup 5 --dtype $dt --no-e2e --no-cpu --no-train --no-alt 2>> gpurun_out/r02_s26_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['sustained_100_steps']['value'], d['clocks']['sm_mhz'], d['launches_per_step'])" >> gpurun_out/r02_s26_ab.log
  done
done
cat gpurun_out/r02_s26_ab.log; tail -3 gpurun_out/r02_s26_ab.err
