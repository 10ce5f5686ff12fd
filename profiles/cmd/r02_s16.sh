# round 2, call 16: schedule sweep with the fused prologue -- iteration batch and mask-build fork point
for v in "B2U_MC_FORK=enc0 NB=8" "B2U_MC_FORK=enc0 NB=10" "B2U_MC_FORK=enc0 NB=12" "B2U_MC_FORK=enc0 NB=16" "B2U_MC_FORK=enc1 NB=10" "B2U_MC_FORK=enc2 NB=10" "B2U_MC_FORK=bottleneck NB=10" "B2U_MC_FORK=dec1 NB=10"; do
  nb=$(echo $v | sed 's/.*NB=//')
  echo "== $v" >> gpurun_out/r02_s16_ab.log
  env $v python bench.py --steps 40 --warmup 5 --iter-batch $nb --no-e2e --no-cpu --no-train --no-alt 2>> gpurun_out/r02_s16_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['sustained_100_steps']['value'], d['clocks']['sm_mhz'])" >> gpurun_out/r02_s16_ab.log
done
cat gpurun_out/r02_s16_ab.log; tail -3 gpurun_out/r02_s16_ab.err
