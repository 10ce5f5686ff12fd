# round 2, call 35: ncu source view of the plain 64 -> 64 level-0 conv after the combined statistics reduction
ncu --set full --clock-control none --import-source on -k regex:conv3x3_v2 -s 3 -c 1 -f -o gpurun_out/r02_s35_conv64 python tests/prof_conv.py 592 576 64 64 10 > gpurun_out/r02_s35_ncu.log 2>&1
tail -2 gpurun_out/r02_s35_ncu.log
python tests/gpu_diag.py convbench 2>&1 | tail -25
