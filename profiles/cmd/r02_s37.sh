# round 2, call 37: software-pipelined TMEM loads in the MT = 2 epilogue of the plain conv kernels; MT 1 for 64 -> 128
python -m pytest tests/test_gpu_parity.py -x -q -k "conv3x3 or fused or forward or mc_dropblock or mc_full or remainder or train_step or backward" 2>&1 | tail -2
python tests/exp_conv_plan.py 10 fp16 2>&1 | cut -c1-40,150-200
for rep in 1 2; do
for v in "B2U_FUSED=1" "B2U_FUSED=0"; do
  for dt in fp16; do
  echo "== $v $dt" >> gpurun_out/r02_s37_ab.log
  env $v python bench.py --steps 40 --warmup 5 --dtype $dt --no-e2e --no-cpu --no-train --no-alt 2>> gpurun_out/r02_s37_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['conv_ms_per_step'], d['roofline']['sustained_100_steps']['value'], d['clocks']['sm_mhz'])" >> gpurun_out/r02_s37_ab.log
  done
done
done
cat gpurun_out/r02_s37_ab.log; tail -3 gpurun_out/r02_s37_ab.err
