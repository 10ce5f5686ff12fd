# round 2, call 74 (--gpus 2): multi-GPU correctness pytest + bench at N = 2 on the r02i state (final tree) (smoke + reference arm first)
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q -k "multi_gpu" > gpurun_out/r02i_mg2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02i_mg2_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 --no-cpu --no-libbar > gpurun_out/r02i_mg2_bench.json 2> gpurun_out/r02i_mg2_bench.err; echo "bench n2 rc=$?"
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r02i_mg2_bench.json') if l.startswith('{')][-1])
e=d['e2e']; print(d['value'], d['n_gpus'], d['ms_per_step'], e['value'], e['seconds_per_call'], e['checksum'], e['rotation_ensemble']['value'], e['train']['value'], e['train']['ms_per_step'], e.get('allreduce_ms'))
P
