python -m pytest tests -m gpu -x -q > gpurun_out/s15_pytest.log 2>&1; tail -3 gpurun_out/s15_pytest.log
python tests/exp_overlap.py 10 > gpurun_out/s15_overlap.log 2>&1; cat gpurun_out/s15_overlap.log
