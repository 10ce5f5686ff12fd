python -m pytest tests -m gpu -x -q > gpurun_out/s14_pytest.log 2>&1; tail -3 gpurun_out/s14_pytest.log
for k in 1 2 3 4 8; do
echo "== B2U_MASK_BLOCKS_PER_SM=$k"
B2U_MASK_BLOCKS_PER_SM=$k python tests/exp_overlap.py 10 2>&1 | head -4
done > gpurun_out/s14_overlap.log 2>&1; cat gpurun_out/s14_overlap.log
