# round 2, call 77 (the last GPU seconds of the round): home-made initcheck -- the eager MC call with every workspace / mask buffer poisoned (0xFF) before the second run
timeout 24 python -u tests/exp_poison.py > gpurun_out/r02j_poison.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r02j_poison.log
