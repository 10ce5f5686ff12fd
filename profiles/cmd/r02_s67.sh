# round 2, call 67: final tree (backward kernels specialised): smoke, full GPU suite, the driver's default bench command
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r02h_pytest.log
( time python bench.py > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err ) 2>&1 | grep real
python -c 'import json; d=json.load(open("gpurun_out/r02h_bench.json")); e=d["e2e"]; print(d["value"], d["ms_per_step"], e["value"], e["seconds_per_call"], e["checksum"]["samples_sha256"], e["rotation_ensemble"]["value"], e["train"]["value"], e["train"]["ms_per_step"], d["roofline"]["frac"], d["roofline"]["train"]["whole_step"])'
