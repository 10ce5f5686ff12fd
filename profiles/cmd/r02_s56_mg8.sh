# round 2, call 56 (--gpus 8): multi-GPU correctness pytest over 8 ranks + bench at N = 8 on the r02d state
timeout 900 python -m pytest tests -m gpu -x -q -k "multi_gpu" > gpurun_out/r02d_mg8_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02d_mg8_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 30 --warmup 3 --no-cpu --no-libbar --no-sweep > gpurun_out/r02d_mg8_bench.json 2> gpurun_out/r02d_mg8_bench.err; echo "bench n8 rc=$?"
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r02d_mg8_bench.json') if l.startswith('{')][-1])
e=d['e2e']; print(d['value'], d['n_gpus'], d['ms_per_step'], e['value'], e['seconds_per_call'], e['checksum'], e['rotation_ensemble']['value'], e['rotation_ensemble']['seconds_per_call'], e['train']['value'], e['train']['ms_per_step'], e.get('allreduce_ms'))
P
