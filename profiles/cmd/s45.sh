python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dropblock or mc_ or train_step or ichan" > gpurun_out/s45_pytest.log 2>&1; tail -2 gpurun_out/s45_pytest.log
python tests/exp_overlap.py 10 enc0 2>&1 | head -6
