# round 2, call 66: backward elementwise kernels specialised by gradient source (A / A+pool / head): parity, train step
timeout 900 python -m pytest tests -m gpu -x -q -k "train or bwd or backward or gradient or reference_own" > gpurun_out/s66_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/s66_pytest.log
for rep in 1 2; do timeout 300 python tests/exp_train_skip.py 40 2>&1 | tail -1; done | tee gpurun_out/s66_train.log
B2U_EXP_SKIP_CALLS=b2u_unit_bwd_stats,b2u_unit_bwd_apply timeout 300 python tests/exp_train_skip.py 40 2>&1 | tail -1 | tee -a gpurun_out/s66_train.log
