for k in 0 4096 8192 16384 32768 65536; do
echo "== B2U_MASK_SMEM=$k"
B2U_MASK_SMEM=$k python tests/exp_overlap.py 10 enc0 2>&1 | head -4
done > gpurun_out/s39_overlap.log 2>&1; cat gpurun_out/s39_overlap.log
