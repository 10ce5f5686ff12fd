# round 2, call 73: r02i evidence on the final tree (head kernel = cp.async ring, wgrad TMA stores): full GPU suite, the driver's default bench
# command, head A/B with the library defaults, per-launch ncu metrics of one eager training step, ncu time of the head in one MC step
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests -m gpu -x -q > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r02i_pytest.log
python bench.py > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02i_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["seconds_per_call"], d["e2e"]["checksum"]["samples_sha256"],
      d["e2e"]["rotation_ensemble"]["value"], d["train"]["value"], d["train"]["ms_per_step"], d["roofline"]["frac"],
      {k: round(v["frac_of_hbm_peak"], 3) for k, v in d["roofline"]["hbm_bound_kernels"].items()}, d["clocks"])
PY
timeout 120 python tests/exp_head.py 10 2>&1 | grep -E "us |exp_head" | head -12 | tee gpurun_out/r02i_head.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size
timeout 240 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02i_train_metrics.csv python tests/prof_train.py > gpurun_out/r02i_ncu_train.log 2>&1; echo "ncu train rc=$?"; tail -1 gpurun_out/r02i_ncu_train.log
timeout 120 ncu --metrics $M --clock-control none --profile-from-start off -k regex:head8 --csv --log-file gpurun_out/r02i_head_metrics.csv python tests/prof_step.py 10 2 > gpurun_out/r02i_ncu_head.log 2>&1; echo "ncu head rc=$?"; grep -c head8 gpurun_out/r02i_head_metrics.csv
