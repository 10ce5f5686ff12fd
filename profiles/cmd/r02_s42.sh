# round 2, call 42: which levels should fuse now that the plain conv epilogue uses TMA stores (64->64: 281 -> 241 us plain, fused 545)?
# + gradient equal-precision calibration + MC tests after the overlapped remainder prologue
L=$PWD/unet_research_b200/csrc
B="python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu --no-train --no-alt --no-libbar"
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"])'
: > gpurun_out/s42_ab.log
for rep in 1 2; do
for lib in libb2u.so libb2u_tma256.so; do
for lv in "0,1,2,3,4" "1,2,3,4" "2,3,4" ""; do
  echo "== bench $lib fuse_levels=[$lv]" >> gpurun_out/s42_ab.log
  B2U_LIB=$L/$lib B2U_FUSE_LEVELS="$lv" timeout 300 $B 2>/dev/null | python -c "$P" >> gpurun_out/s42_ab.log 2>&1
done; done; done
cat gpurun_out/s42_ab.log
B2U_VERBOSE=1 timeout 600 python tests/gpu_diag.py gradprec > gpurun_out/s42_gradprec.log 2>&1; tail -90 gpurun_out/s42_gradprec.log
timeout 900 python -m pytest tests -m gpu -x -q -k "mc or MC or monte or rotation or reference_own or multi" > gpurun_out/s42_pytest.log 2>&1; tail -5 gpurun_out/s42_pytest.log
