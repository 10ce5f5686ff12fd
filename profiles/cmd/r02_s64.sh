# round 2, call 64: training mask build on a side stream next to the weight repack: train parity tests, MC tests after the plan_segments refactor, A/B
timeout 900 python -m pytest tests -m gpu -x -q -k "train or mc or MC or remainder or reference_own" > gpurun_out/s64_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/s64_pytest.log
for rep in 1 2 3; do
  timeout 300 python tests/exp_train_skip.py 40 2>&1 | tail -1
  B2U_TRAIN_MASK_OVERLAP=0 timeout 300 python tests/exp_train_skip.py 40 2>&1 | tail -1 | sed 's/$/ (mask build in line)/'
done | tee gpurun_out/s64_train.log
