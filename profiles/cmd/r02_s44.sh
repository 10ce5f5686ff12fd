# round 2, call 44: new defaults (TMA stores in every 16-bit conv kernel, levels 1..4 fused, level 0 two-pass): full GPU suite + bench + gradient calibration
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s44_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s44_pytest.log; tail -4 gpurun_out/s44_pytest.log
B2U_VERBOSE=1 timeout 600 python tests/gpu_diag.py gradprec > gpurun_out/s44_gradprec.log 2>&1; grep -v "^      " gpurun_out/s44_gradprec.log | tail -5
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/s44_bench.json 2> gpurun_out/s44_bench.err; echo "bench rc=$?"; head -c 600 gpurun_out/s44_bench.json
