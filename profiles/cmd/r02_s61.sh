# round 2, call 61: final state after the reverted experiments -- smoke, full GPU suite twice, short bench
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
for i in 1 2; do timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02f_pytest_$i.log 2>&1; echo "run $i rc=$?"; tail -1 gpurun_out/r02f_pytest_$i.log; grep -E "^FAILED" gpurun_out/r02f_pytest_$i.log; done
python bench.py --steps 20 --warmup 5 --no-cpu --no-libbar 2>/dev/null | python -c 'import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["checksum"]["samples_sha256"], d["e2e"]["train"]["value"])'
