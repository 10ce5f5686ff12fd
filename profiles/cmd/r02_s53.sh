# round 2, call 53: operand-swapped weight gradient for Cg = 64 (nine taps in one CTA, M full): parity, stand-alone timing, train step A/B
timeout 600 python tests/gpu_diag.py wgrad > gpurun_out/s53_wgrad.log 2>&1; tail -12 gpurun_out/s53_wgrad.log
timeout 900 python -m pytest tests -m gpu -x -q -k "wgrad or train or bwd or backward" > gpurun_out/s53_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s53_pytest.log
for rep in 1 2; do
  timeout 300 python tests/exp_train_skip.py 40 2>&1 | tail -1
  B2U_WGRAD_SWAP64=0 timeout 300 python tests/exp_train_skip.py 40 2>&1 | tail -1
done | tee gpurun_out/s53_train.log
