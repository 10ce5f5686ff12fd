# round 2, call 59: fork-point sweep of the mask build on the r02d kernels (r02b: enc0 best)
B="python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu --no-train --no-alt --no-libbar"
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"])'
: > gpurun_out/s59_fork.log
for rep in 1 2; do for fk in enc0 enc1 enc2 enc3 bottleneck; do echo "fork $fk" >> gpurun_out/s59_fork.log; B2U_MC_FORK=$fk timeout 300 $B 2>/dev/null | python -c "$P" >> gpurun_out/s59_fork.log 2>&1; done; done
cat gpurun_out/s59_fork.log
