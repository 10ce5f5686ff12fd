# round 2, call 10: fused prologue with a deeper patch ring; fp16 packing through F2FP.SATFINITE
python tests/gpu_diag.py convpro 2>&1 | grep -c "identical True"
python tests/exp_convpro.py 10 > gpurun_out/r02_s10_exp.log 2>&1; cat gpurun_out/r02_s10_exp.log
for v in "B2U_FUSED=0" "B2U_FUSED=1"; do
  for dt in bf16 fp16; do
  echo "== $v $dt" >> gpurun_out/r02_s10_ab.log
  env $v python bench.py --steps 40 --warmup 5 --dtype $dt --no-e2e --no-cpu --no-train --no-alt 2>> gpurun_out/r02_s10_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['conv_ms_per_step'], d['roofline']['sustained_100_steps']['value'], d['clocks']['sm_mhz'], json.dumps(d['roofline']['other_kernels_ms_per_step']))" >> gpurun_out/r02_s10_ab.log
  done
done
cat gpurun_out/r02_s10_ab.log | cut -c1-130; tail -5 gpurun_out/r02_s10_ab.err
python -m pytest tests/test_gpu_parity.py -x -q -k "fp16 or precision or forward_default" 2>&1 | tail -3
