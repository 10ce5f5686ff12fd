python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "conv3x3 or forward_bf16 or tf32 or train_step" > gpurun_out/s40_pytest.log 2>&1; tail -2 gpurun_out/s40_pytest.log
python tests/exp_overlap.py 10 enc0 2>&1 | head -8
python bench.py --steps 30 --no-cpu --no-e2e > gpurun_out/s40_bench.json 2> gpurun_out/s40_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/s40_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['train']['value'], d['train']['ms_per_step'], d['clocks'])
P
