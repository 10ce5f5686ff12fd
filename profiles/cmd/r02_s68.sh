# round 2, call 68: TMA-store epilogue of the weight-gradient kernel (32 x 32 fp32 blocks staged in the idle pipeline stages): parity, A/B
timeout 300 python tests/gpu_diag.py wgrad > gpurun_out/s68_wgrad.log 2>&1; tail -11 gpurun_out/s68_wgrad.log
timeout 900 python -m pytest tests -m gpu -x -q -k "train or wgrad or bwd or backward or gradient" > gpurun_out/s68_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/s68_pytest.log
for rep in 1 2; do
  timeout 300 python tests/exp_train_skip.py 40 2>&1 | tail -1
  B2U_WGRAD_TMA_STORE=0 timeout 300 python tests/exp_train_skip.py 40 2>&1 | tail -1 | sed 's/$/ (wgrad per-thread stores)/'
done | tee gpurun_out/s68_train.log
