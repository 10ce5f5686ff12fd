# round 2, call 38: same-box A/B: pipelined TMEM loads in the plain MT = 2 epilogue (new) vs the previous build (old)
for lib in "" "/root/repo/unet_research_b200/csrc/libb2u_old.so"; do
  echo "== lib=$lib" >> gpurun_out/r02_s38_plan.log
  B2U_LIB=$lib python tests/exp_conv_plan.py 10 fp16 2>&1 | grep -E "plain|64->  64|128-> 128" | sed 's/  */ /g' | cut -c1-230 >> gpurun_out/r02_s38_plan.log
done
cat gpurun_out/r02_s38_plan.log
for rep in 1 2; do
for lib in "" "/root/repo/unet_research_b200/csrc/libb2u_old.so"; do
  echo "== lib=$lib" >> gpurun_out/r02_s38_ab.log
  B2U_LIB=$lib python bench.py --steps 40 --warmup 5 --no-e2e --no-cpu --no-train --no-alt 2>> gpurun_out/r02_s38_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['conv_ms_per_step'], d['roofline']['sustained_100_steps']['value'], d['clocks']['sm_mhz'])" >> gpurun_out/r02_s38_ab.log
done
done
cat gpurun_out/r02_s38_ab.log; tail -3 gpurun_out/r02_s38_ab.err
