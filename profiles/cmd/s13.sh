python -m pytest tests -m gpu -x -q > gpurun_out/s13_pytest.log 2>&1; tail -3 gpurun_out/s13_pytest.log
python bench.py --steps 30 --no-cpu --no-train --no-e2e > gpurun_out/s13_bench.json 2> gpurun_out/s13_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/s13_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['other_kernels_ms_per_step'])
P
python tests/exp_overlap.py 10 > gpurun_out/s13_overlap.log 2>&1; cat gpurun_out/s13_overlap.log
