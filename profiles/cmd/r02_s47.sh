# round 2, call 47: under-filled-grid tile selection (batch-1 deep layers) + F16 convT dgrad pack fix: parity, train step, batch-1 sweep
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s47_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s47_pytest.log; tail -4 gpurun_out/s47_pytest.log
for rep in 1 2; do timeout 300 python tests/exp_train_skip.py 40 2>&1 | tail -1; done | tee gpurun_out/s47_train.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-libbar > gpurun_out/s47_bench.json 2> gpurun_out/s47_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/s47_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['seconds_per_call'], d['e2e']['rotation_ensemble']['value'])
print([ (s['size'], round(s['imgs_per_s'])) for s in d['e2e']['sweep']])
print(d['e2e']['train']['value'], d['e2e']['train']['ms_per_step'])
P
