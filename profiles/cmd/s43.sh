python bench.py --steps 30 --no-cpu --no-train > gpurun_out/s43_bench.json 2> gpurun_out/s43_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/s43_bench.json')); print(d['value'], d['ms_per_step'], d['e2e'])
P
