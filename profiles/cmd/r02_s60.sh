# round 2, call 60: two-phase pipelined mask build (dilation first, next to the encoder; centres for the step after): MC parity tests, A/B
timeout 900 python -m pytest tests -m gpu -x -q -k "mc or MC or fused_and_unfused or ichan or reference_own or dilate" > gpurun_out/s60_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s60_pytest.log
B="python bench.py --steps 30 --warmup 5 --no-cpu --no-train --no-alt --no-libbar --no-rotation --no-sweep"
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], "e2e", d["e2e"]["value"], d["e2e"]["seconds_per_call"], d["e2e"]["checksum"]["samples_sha256"])'
: > gpurun_out/s60_ab.log
for rep in 1 2 3; do
  echo "pipeline on" >> gpurun_out/s60_ab.log; timeout 300 $B 2>/dev/null | python -c "$P" >> gpurun_out/s60_ab.log 2>&1
  echo "pipeline off" >> gpurun_out/s60_ab.log; B2U_MC_PIPELINE=0 timeout 300 $B 2>/dev/null | python -c "$P" >> gpurun_out/s60_ab.log 2>&1
done
echo "pipeline on, fork enc0" >> gpurun_out/s60_ab.log; B2U_MC_FORK=enc0 timeout 300 $B 2>/dev/null | python -c "$P" >> gpurun_out/s60_ab.log 2>&1
echo "pipeline on, fork enc2" >> gpurun_out/s60_ab.log; B2U_MC_FORK=enc2 timeout 300 $B 2>/dev/null | python -c "$P" >> gpurun_out/s60_ab.log 2>&1
cat gpurun_out/s60_ab.log
