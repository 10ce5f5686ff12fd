python tests/exp_overlap.py 10 enc0 2>&1 | tee gpurun_out/s35_overlap.log
