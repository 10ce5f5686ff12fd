# round 2, call 31: one GroupNorm-statistics reduction per MT = 2 item instead of per tile (conv3x3 epilogue)
python -m pytest tests/test_gpu_parity.py -x -q -k "conv3x3 or fused or forward or mc_dropblock or mc_full or remainder or train_step" 2>&1 | tail -2
for c in 0 1; do echo "== B2U_COMBINE_STATS=$c" >> gpurun_out/r02_s31_exp.log; B2U_COMBINE_STATS=$c python tests/exp_convpro.py 10 fp16 2>&1 | head -3 >> gpurun_out/r02_s31_exp.log; done; cat gpurun_out/r02_s31_exp.log
for rep in 1 2; do
for v in "B2U_COMBINE_STATS=0" "B2U_COMBINE_STATS=1"; do
  for dt in fp16; do
  echo "== $v $dt" >> gpurun_out/r02_s31_ab.log
  env $v python bench.py --steps 40 --warmup 5 --dtype $dt --no-e2e --no-cpu --no-train --no-alt 2>> gpurun_out/r02_s31_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['conv_ms_per_step'], d['roofline']['sustained_100_steps']['value'], d['clocks']['sm_mhz'])" >> gpurun_out/r02_s31_ab.log
  done
done
done
cat gpurun_out/r02_s31_ab.log; tail -3 gpurun_out/r02_s31_ab.err
