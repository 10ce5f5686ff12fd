# round 2, call 43: upper-bound experiments on the captured TRAINING step (timing only; skipped kernels leave stale data)
: > gpurun_out/s43_train_skip.log
run() { B2U_EXP_SKIP_CALLS="$1" timeout 300 python tests/exp_train_skip.py 40 2>&1 | tail -1 >> gpurun_out/s43_train_skip.log; }
for rep in 1 2; do
run ""
run "b2u_pack_conv3x3_weight_pair,b2u_pack_convT2x2_weight,b2u_pack_convT2x2_dgrad_weight"
run "b2u_unit_bwd_finalize"
run "b2u_gn_finalize_ex,b2u_gn_finalize"
run "b2u_wgrad,b2u_wgrad_first"
run "b2u_dropblock_centers,b2u_dropblock_dilate_v2"
run "b2u_unit_bwd_stats,b2u_unit_bwd_apply"
run "b2u_sgd_step"
B2U_EXP_SKIP_WGRAD_REDUCE=1 timeout 300 python tests/exp_train_skip.py 40 2>&1 | tail -1 | sed 's/$/  (wgrad split-K reduce skipped)/' >> gpurun_out/s43_train_skip.log
done
cat gpurun_out/s43_train_skip.log
timeout 600 python -m pytest tests -m gpu -x -q -k "train or wgrad or first" > gpurun_out/s43_pytest.log 2>&1; tail -3 gpurun_out/s43_pytest.log
