# round 2, call 4: fused conv prologue v3 (explicit LDS/STS) -- per-layer timing + ncu source-level capture of the level-0 kernel
python tests/gpu_diag.py convpro 2>&1 | grep -c "identical True" > gpurun_out/r02_s04_convpro.log
python tests/exp_convpro.py 10 > gpurun_out/r02_s04_exp.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv3x3_v2 -s 2 -c 1 -f -o gpurun_out/r02_s04_pro64 python tests/prof_convpro.py > gpurun_out/r02_s04_ncu.log 2>&1
cat gpurun_out/r02_s04_convpro.log; cat gpurun_out/r02_s04_exp.log; tail -3 gpurun_out/r02_s04_ncu.log
