python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/s38_bench_n2.json 2> gpurun_out/s38_bench_n2.err; echo "bench n2 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/multigpu_check.py > gpurun_out/s38_mg2.log 2>&1; echo "mgcheck rc=$?"; tail -8 gpurun_out/s38_mg2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/s38_ref_n2.json 2> gpurun_out/s38_ref_n2.err; echo "ref rc=$?"; cat gpurun_out/s38_ref_n2.json | cut -c1-600
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/s38_bench_n2.json') if l.startswith('{')][-1]); print(d['value'], d['n_gpus'], d['ms_per_step'], d['e2e'], d['train']['value'], d['train'].get('allreduce'), d.get('allreduce_ms'))
P
