# round 2, call 30: same-box A/B of the two fused-prologue builds (4 transform warps at 128 registers vs 8 at 80)
for rep in 1 2; do
for v in "B2U_LIB=" "B2U_LIB=/root/repo/unet_research_b200/csrc/libb2u_8w.so"; do
  for dt in fp16; do
  echo "== $v $dt" >> gpurun_out/r02_s30_ab.log
  env $v python bench.py --steps 40 --warmup 5 --dtype $dt --no-e2e --no-cpu --no-train --no-alt 2>> gpurun_out/r02_s30_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['conv_ms_per_step'], d['roofline']['sustained_100_steps']['value'], d['clocks']['sm_mhz'])" >> gpurun_out/r02_s30_ab.log
  done
done
done
cat gpurun_out/r02_s30_ab.log; tail -3 gpurun_out/r02_s30_ab.err
