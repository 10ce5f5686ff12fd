# round 2, call 11: timeline of the overlapped MC step (events on every launch, both streams)
python tests/exp_timeline.py 10 bf16 > gpurun_out/r02_s11_timeline.log 2>&1
B2U_FUSED=0 python tests/exp_timeline.py 10 bf16 > gpurun_out/r02_s11_timeline_unfused.log 2>&1
head -60 gpurun_out/r02_s11_timeline.log
