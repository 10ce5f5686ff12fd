python tests/exp_conv64.py 10 2>&1 | tee gpurun_out/s33_conv64.log
