# round 2, call 24: timeline after the v2 dilation + finalize change
python tests/exp_timeline.py 10 fp16 > gpurun_out/r02_s24_timeline.log 2>&1
head -24 gpurun_out/r02_s24_timeline.log
