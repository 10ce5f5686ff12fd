# round 2, call 17: v2 dilation (sparse NHWC scatter + word-parallel 7x7 OR) -- parity and step A/B
python -m pytest tests/test_gpu_parity.py -x -q -k "dropblock or dilate or ichan or mc_dropblock_vs" > gpurun_out/r02_s17_pytest.log 2>&1; tail -4 gpurun_out/r02_s17_pytest.log
for v in "B2U_DILATE=v1" "B2U_DILATE=v2"; do
  for dt in bf16 fp16; do
  echo "== $v $dt" >> gpurun_out/r02_s17_ab.log
  env $v python bench.py --steps 40 --warmup 5 --dtype $dt --no-e2e --no-cpu --no-train --no-alt 2>> gpurun_out/r02_s17_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
o=d['roofline']['other_kernels_ms_per_step']
print(d['value'], d['ms_per_step'], d['roofline']['sustained_100_steps']['value'], d['clocks']['sm_mhz'], {k:round(v,3) for k,v in o.items() if 'dropblock' in k})" >> gpurun_out/r02_s17_ab.log
  done
done
cat gpurun_out/r02_s17_ab.log; tail -3 gpurun_out/r02_s17_ab.err
python tests/exp_timeline.py 10 bf16 2>&1 | head -12
