# round 2, call 5: fused conv prologue v4 (fast division, interior shortcut, batched loads, predicated stores) -- per-layer timing + ncu source-level capture of the level-0 kernel
python tests/gpu_diag.py convpro 2>&1 | grep -c "identical True" > gpurun_out/r02_s05_convpro.log
python tests/exp_convpro.py 10 > gpurun_out/r02_s05_exp.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv3x3_v2 -s 2 -c 1 -f -o gpurun_out/r02_s05_pro64 python tests/prof_convpro.py > gpurun_out/r02_s05_ncu.log 2>&1
cat gpurun_out/r02_s05_convpro.log; cat gpurun_out/r02_s05_exp.log; tail -3 gpurun_out/r02_s05_ncu.log
