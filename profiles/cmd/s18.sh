python -m pytest tests -m gpu -x -q > gpurun_out/s18_pytest.log 2>&1; tail -3 gpurun_out/s18_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 30 --no-cpu --no-e2e --no-train > gpurun_out/s18_bench.json 2> gpurun_out/s18_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/s18_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['other_kernels_ms_per_step'], d['roofline']['traffic'])
P
