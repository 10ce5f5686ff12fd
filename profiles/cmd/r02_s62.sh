# round 2, call 62: bench with the training size sweep (configs[4] training part)
( time python bench.py --steps 20 --warmup 5 --no-cpu --no-libbar > gpurun_out/s62_bench.json 2> gpurun_out/s62_bench.err ) 2>&1 | grep real; tail -3 gpurun_out/s62_bench.err
python -c 'import json; d=json.load(open("gpurun_out/s62_bench.json")); print(d["value"], d["e2e"]["train"])'
