# round 2, call 14: fused masked BCE, 3-bucket DDP backward graphs -- tests + train bench
python -m pytest tests/test_gpu_parity.py -x -q -k "bce or train or reference_own or gradient_acc" > gpurun_out/r02_s14_pytest.log 2>&1; tail -4 gpurun_out/r02_s14_pytest.log
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-alt 2> gpurun_out/r02_s14.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
t=d['train']; print(t['value'], t['ms_per_step'], t['loss_first'], t['loss_last']); print(json.dumps(t['roofline']['elementwise_ms'])); print(json.dumps(t['roofline']['families']))"
tail -3 gpurun_out/r02_s14.err
