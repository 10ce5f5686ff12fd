# round 2, call 9: does capping the occupancy of gn_apply (registers left for the overlapped mask build) help the step?
for v in "B2U_FUSED=0 B2U_APPLY_SMEM_KB=0" "B2U_FUSED=0 B2U_APPLY_SMEM_KB=72" "B2U_FUSED=0 B2U_APPLY_SMEM_KB=110" "B2U_FUSED=1 B2U_APPLY_SMEM_KB=0" "B2U_FUSED=1 B2U_APPLY_SMEM_KB=72" "B2U_FUSED=1 B2U_APPLY_SMEM_KB=110"; do
  echo "== $v" >> gpurun_out/r02_s09_ab.log
  env $v python bench.py --steps 40 --warmup 5 --dtype bf16 --no-e2e --no-cpu --no-train --no-alt 2>> gpurun_out/r02_s09_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['conv_ms_per_step'], d['roofline']['sustained_100_steps']['value'], d['clocks']['sm_mhz'], json.dumps(d['roofline']['other_kernels_ms_per_step']))" >> gpurun_out/r02_s09_ab.log
done
cat gpurun_out/r02_s09_ab.log; tail -5 gpurun_out/r02_s09_ab.err
