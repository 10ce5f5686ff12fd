# round 2, call 48: under-filled-grid BLOCK_N selection with MT kept (batch-independent statistics rounding): full suite + train step
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s48_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s48_pytest.log; tail -4 gpurun_out/s48_pytest.log
for rep in 1 2; do timeout 300 python tests/exp_train_skip.py 40 2>&1 | tail -1; done | tee gpurun_out/s48_train.log
