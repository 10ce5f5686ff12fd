# round 2, call 72: head kernel -- cp.async ring variant and block sizes 256 / 224 / 192 / 160 (fill of the last trip): bit-compare, A/B
timeout 300 python tests/exp_head.py 20 > gpurun_out/s72_head.log 2>&1; echo "exp_head rc=$?"; tail -22 gpurun_out/s72_head.log
timeout 200 python tests/gpu_diag.py head 2>&1 | tail -3
