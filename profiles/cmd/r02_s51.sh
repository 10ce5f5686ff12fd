# round 2, call 51: flake hunt -- the full GPU suite four times (no -x), train step after the reduce9 block-shape fix
for i in 1 2 3 4; do timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/s51_pytest_$i.log 2>&1; echo "run $i rc=$?"; tail -3 gpurun_out/s51_pytest_$i.log | head -2; grep -E "^FAILED" gpurun_out/s51_pytest_$i.log; done
for rep in 1 2; do timeout 300 python tests/exp_train_skip.py 40 2>&1 | tail -1; done | tee gpurun_out/s51_train.log
