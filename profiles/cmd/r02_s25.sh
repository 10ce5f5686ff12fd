# round 2, call 25: upper bound of folding gn_finalize away (timing only: coefficients are stale after warm-up)
for v in "B2U_SKIP_FINALIZE=" "B2U_SKIP_FINALIZE=conv" "B2U_SKIP_FINALIZE=all"; do
  for dt in fp16; do
  echo "== $v $dt" >> gpurun_out/r02_s25_ab.log
  env $v python bench.py --steps 40 --warmup 5 --dtype $dt --no-e2e --no-cpu --no-train --no-alt 2>> gpurun_out/r02_s25_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['sustained_100_steps']['value'], d['clocks']['sm_mhz'], d['launches_per_step'])" >> gpurun_out/r02_s25_ab.log
  done
done
cat gpurun_out/r02_s25_ab.log; tail -3 gpurun_out/r02_s25_ab.err
