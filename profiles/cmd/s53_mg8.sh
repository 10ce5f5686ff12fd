python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 30 --warmup 3 --no-cpu > gpurun_out/s42_bench_n8.json 2> gpurun_out/s42_bench_n8.err; echo "rc=$?"
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/s42_bench_n8.json') if l.startswith('{')][-1]); print(d['value'], d['n_gpus'], d['ms_per_step'], d['e2e']['value'], d['train']['value'], d.get('allreduce_ms'))
P
