# round 2, call 36: tile-plan sweep at batch 10 (plain and fused kernels)
python tests/exp_conv_plan.py 10 fp16 > gpurun_out/r02_s36_plan.log 2>&1; cat gpurun_out/r02_s36_plan.log
