python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "conv3x3 or forward_bf16 or tf32" > gpurun_out/s28_pytest.log 2>&1; tail -3 gpurun_out/s28_pytest.log
python bench.py --steps 30 --no-cpu --no-e2e --no-train > gpurun_out/s28_bench.json 2> gpurun_out/s28_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/s28_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['conv_ms_per_step'], d['clocks'])
P
