python tests/gpu_diag.py libbar 2>&1 | tee gpurun_out/s47_libbar.log | tail -12
