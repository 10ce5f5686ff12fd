python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
( time python bench.py > gpurun_out/s44_bench.json 2> gpurun_out/s44_bench.err ) 2>&1 | grep real
( time python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/s44_ref.json 2> gpurun_out/s44_ref.err ) 2>&1 | grep real
wc -l gpurun_out/s44_bench.json gpurun_out/s44_ref.json; cut -c1-400 gpurun_out/s44_ref.json
python - <<'P'
import json
d=json.load(open('gpurun_out/s44_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['train']['value'], d['cpu_baseline'])
P
