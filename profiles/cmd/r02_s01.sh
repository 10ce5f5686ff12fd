# round 2, call 1: full GPU test-suite + bench line (fp16 default) after the parity / rotation / bench changes
python -m pytest tests -m gpu -x -q > gpurun_out/r02_s01_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_s01_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_s01_bench.json 2> gpurun_out/r02_s01_bench.err; echo "bench rc=$?" >> gpurun_out/r02_s01_bench.err
tail -5 gpurun_out/r02_s01_pytest.log; tail -3 gpurun_out/r02_s01_bench.err; head -c 600 gpurun_out/r02_s01_bench.json
