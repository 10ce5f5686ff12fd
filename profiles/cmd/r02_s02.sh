# round 2, call 2: fused conv prologue -- parity, per-layer timing, step A/B
python tests/gpu_diag.py convpro > gpurun_out/r02_s02_convpro.log 2>&1
python tests/exp_convpro.py 10 > gpurun_out/r02_s02_exp.log 2>&1
python tests/exp_convpro.py 10 fp16 >> gpurun_out/r02_s02_exp.log 2>&1
for v in "B2U_FUSED=0" "B2U_FUSED=1" "B2U_FUSE_LEVELS=1,2,3,4"; do
  for dt in bf16 fp16; do
    echo "== $v $dt" >> gpurun_out/r02_s02_ab.log
    env $v python bench.py --steps 40 --warmup 5 --dtype $dt --no-e2e --no-cpu --no-train --no-alt 2>> gpurun_out/r02_s02_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['conv_ms_per_step'], d['roofline']['sustained_100_steps']['value'], d['clocks']['sm_mhz'], json.dumps(d['roofline']['other_kernels_ms_per_step']))" >> gpurun_out/r02_s02_ab.log
  done
done
python -m pytest tests/test_gpu_parity.py -x -q -k "fused or forward_default or mc_dropblock" > gpurun_out/r02_s02_pytest.log 2>&1
cat gpurun_out/r02_s02_convpro.log | tail -20; cat gpurun_out/r02_s02_exp.log; cat gpurun_out/r02_s02_ab.log; tail -5 gpurun_out/r02_s02_pytest.log; tail -5 gpurun_out/r02_s02_ab.err
