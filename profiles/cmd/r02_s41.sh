# round 2, call 41: TMA-store epilogues (convT 5-D pixel-shuffle box; conv3x3 32-channel runs) -- parity, then A/B on one box:
# libb2u.so (conv3x3 TMA stores for BLOCK_N <= 128, convT TMA stores) / libb2u_stg.so (conv3x3 per-thread stores) / libb2u_tma256.so (all)
timeout 600 python -m pytest tests -m gpu -x -q -k "conv or forward or fused or prologue" > gpurun_out/s41_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s41_pytest.log
tail -5 gpurun_out/s41_pytest.log
L=unet_research_b200/csrc
echo "== convT bench: TMA store (default) vs per-thread stores" > gpurun_out/s41_ab.log
timeout 300 python tests/gpu_diag.py convtbench >> gpurun_out/s41_ab.log 2>&1
B2U_CONVT_TMA_STORE=0 timeout 300 python tests/gpu_diag.py convtbench >> gpurun_out/s41_ab.log 2>&1
for lib in libb2u.so libb2u_stg.so libb2u_tma256.so; do
  echo "== exp_convpro fp16 batch 10, $lib" >> gpurun_out/s41_ab.log
  B2U_LIB=$PWD/$L/$lib timeout 300 python tests/exp_convpro.py 10 fp16 >> gpurun_out/s41_ab.log 2>&1
done
for rep in 1 2; do
for lib in libb2u.so libb2u_stg.so libb2u_tma256.so; do
  echo "== bench $lib" >> gpurun_out/s41_ab.log
  B2U_LIB=$PWD/$L/$lib timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu --no-train --no-alt --no-libbar 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['clocks'])" >> gpurun_out/s41_ab.log 2>&1
done
echo "== bench libb2u.so, convT per-thread stores" >> gpurun_out/s41_ab.log
B2U_CONVT_TMA_STORE=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu --no-train --no-alt --no-libbar 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['clocks'])" >> gpurun_out/s41_ab.log 2>&1
done
cat gpurun_out/s41_ab.log
