# round 2, call 54: differentiable square_pad + resize (adjoint kernel) parity
timeout 600 python -m pytest tests -m gpu -x -q -k "resize" > gpurun_out/s54_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/s54_pytest.log
