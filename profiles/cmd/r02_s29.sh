# round 2, call 29: fused prologue with EIGHT transform warps in two groups, 80 registers per thread
python tests/gpu_diag.py convpro 2>&1 | grep -c "identical True"
python tests/exp_convpro.py 10 fp16 > gpurun_out/r02_s29_exp.log 2>&1; cat gpurun_out/r02_s29_exp.log
python -m pytest tests/test_gpu_parity.py -x -q -k "fused or forward_default or mc_dropblock" 2>&1 | tail -2
for dt in bf16 fp16; do
  echo "== fused $dt" >> gpurun_out/r02_s29_ab.log
  python bench.py --steps 40 --warmup 5 --dtype $dt --no-e2e --no-cpu --no-train --no-alt 2>> gpurun_out/r02_s29_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['conv_ms_per_step'], d['roofline']['sustained_100_steps']['value'], d['clocks']['sm_mhz'])" >> gpurun_out/r02_s29_ab.log
done
cat gpurun_out/r02_s29_ab.log; tail -3 gpurun_out/r02_s29_ab.err
