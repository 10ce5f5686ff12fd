# round 2, call 15 (8 GPUs): torchrun-driven multi-GPU pytest + bench at N=8
python -m pytest tests/test_gpu_parity.py -x -q -k multi_gpu > gpurun_out/r02_mg8_pytest.log 2>&1; tail -3 gpurun_out/r02_mg8_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_mg8_bench.json 2> gpurun_out/r02_mg8_bench.err; echo "rc=$?"; tail -3 gpurun_out/r02_mg8_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r02_mg8_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['seconds_per_call'], d['e2e']['checksum'], d['e2e']['rotation_ensemble']['value'], d['train']['value'], d['e2e']['allreduce_ms'])"
