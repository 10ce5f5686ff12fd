# round 2, call 55: conv3x3 tail-wave split + balanced head grid: full suite, then A/B of the MC step and the train step on one box
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s55_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s55_pytest.log
B="python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu --no-train --no-alt --no-libbar"
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["launches_per_step"], d["roofline"]["hbm_bound_kernels"]["b2u_head_fwd"]["frac_of_hbm_peak"], d["roofline"]["frac"])'
: > gpurun_out/s55_ab.log
for rep in 1 2 3; do
  echo "== tail split on" >> gpurun_out/s55_ab.log; timeout 300 $B 2>/dev/null | python -c "$P" >> gpurun_out/s55_ab.log 2>&1
  echo "== tail split off" >> gpurun_out/s55_ab.log; B2U_CONV_TAIL_SPLIT=0 timeout 300 $B 2>/dev/null | python -c "$P" >> gpurun_out/s55_ab.log 2>&1
done
for rep in 1 2; do
  timeout 300 python tests/exp_train_skip.py 40 2>&1 | tail -1 >> gpurun_out/s55_ab.log
  B2U_CONV_TAIL_SPLIT=0 timeout 300 python tests/exp_train_skip.py 40 2>&1 | tail -1 | sed 's/$/ (tail split off)/' >> gpurun_out/s55_ab.log
done
cat gpurun_out/s55_ab.log
