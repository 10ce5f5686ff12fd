# round 2, call 57: iteration-batch sweep on the r02d kernels (the r02b optimum was 8-10)
B="python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu --no-train --no-alt --no-libbar"
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["config"]["iter_batch"], d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"])'
: > gpurun_out/s57_sweep.log
for rep in 1 2; do for ib in 10 8 12 14 16 10; do timeout 300 $B --iter-batch $ib 2>/dev/null | python -c "$P" >> gpurun_out/s57_sweep.log 2>&1; done; done
cat gpurun_out/s57_sweep.log
