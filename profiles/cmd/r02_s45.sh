# round 2, call 45: wgrad tap pairs (N = 128 MMA over two shifted views) + batched weight pack: parity, then A/B of the captured train step
timeout 900 python -m pytest tests -m gpu -x -q -k "wgrad or pack or train or bwd" > gpurun_out/s45_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s45_pytest.log; tail -4 gpurun_out/s45_pytest.log
L=$PWD/unet_research_b200/csrc
: > gpurun_out/s45_train_ab.log
for rep in 1 2; do
  timeout 300 python tests/exp_train_skip.py 40 2>&1 | tail -1 >> gpurun_out/s45_train_ab.log
  B2U_LIB=$L/libb2u_nopair.so timeout 300 python tests/exp_train_skip.py 40 2>&1 | tail -1 >> gpurun_out/s45_train_ab.log
  B2U_BATCHED_PACK=0 timeout 300 python tests/exp_train_skip.py 40 2>&1 | tail -1 | sed 's/$/ (per-tensor pack)/' >> gpurun_out/s45_train_ab.log
done
cat gpurun_out/s45_train_ab.log
timeout 300 python tests/gpu_diag.py wgrad > gpurun_out/s45_wgrad.log 2>&1; tail -30 gpurun_out/s45_wgrad.log
