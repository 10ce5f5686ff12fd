python -m pytest tests -m gpu -x -q > gpurun_out/s37_pytest.log 2>&1; tail -3 gpurun_out/s37_pytest.log
python bench.py --steps 30 --no-cpu --no-e2e > gpurun_out/s37_bench.json 2> gpurun_out/s37_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/s37_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['other_kernels_ms_per_step'], d['train']['value'], d['train']['ms_per_step'], d['clocks'])
P
