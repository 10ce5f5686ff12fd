# round 2, call 75: r02j = the final tree (head launch shape factored into b2u_head_plan, head-variants parity test): the new test alone,
# then the full GPU suite and the driver's default bench command
timeout 120 python tests/gpu_diag.py head_variants 2>&1 | tail -6
python -m pytest tests -m gpu -x -q > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r02j_pytest.log
python bench.py > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02j_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["seconds_per_call"], d["e2e"]["checksum"]["samples_sha256"],
      d["e2e"]["rotation_ensemble"]["value"], d["train"]["value"], d["train"]["ms_per_step"], d["roofline"]["frac"],
      {k: round(v["frac_of_hbm_peak"], 3) for k, v in d["roofline"]["hbm_bound_kernels"].items()}, d["clocks"])
PY
