# round 2, call 63: flake quantification -- the full suite five more times; the once-failed test 25 times in one process after the rotation test
for i in 1 2 3 4 5; do timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/s63_pytest_$i.log 2>&1; echo "run $i rc=$?"; tail -1 gpurun_out/s63_pytest_$i.log; grep -E "^FAILED" gpurun_out/s63_pytest_$i.log; done
python - <<'P' 2>&1 | tail -5
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import test_gpu_parity as T
bad = 0
for i in range(25):
    try:
        if i % 5 == 0:
            T.test_rotation_full_size_properties()
        T.test_mc_full_size_properties()
    except AssertionError as e:
        bad += 1
        print("FAIL at", i, str(e)[:200], flush=True)
print("test_mc_full_size_properties: %d failures in 25 repetitions" % bad)
P
