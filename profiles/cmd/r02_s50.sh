# round 2, call 50: reproducibility hunt -- test_mc_full_size_properties failed once ("reproducible" assert) in the r02d evidence run
L=$PWD/unet_research_b200/csrc
: > gpurun_out/s50_repro.log
run() { echo "== $*" >> gpurun_out/s50_repro.log; env "$@" timeout 300 python tests/exp_repro.py 8 5 10 >> gpurun_out/s50_repro.log 2>&1; }
run B2U_NOP=1
run B2U_NOP=1
run B2U_PDL=0
run B2U_CONVT_TMA_STORE=0
run B2U_LIB=$L/libb2u_stg.so
run B2U_LIB=$L/libb2u_stg.so B2U_CONVT_TMA_STORE=0
run B2U_FUSE_LEVELS=0,1,2,3,4
run B2U_FUSE_LEVELS=
echo "== overlap off" >> gpurun_out/s50_repro.log; timeout 300 python tests/exp_repro.py 8 5 10 0 >> gpurun_out/s50_repro.log 2>&1
echo "== ib 2" >> gpurun_out/s50_repro.log; timeout 300 python tests/exp_repro.py 8 2 10 >> gpurun_out/s50_repro.log 2>&1
cat gpurun_out/s50_repro.log
