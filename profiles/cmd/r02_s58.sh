# round 2, call 58: final state -- smoke, full GPU suite, the driver's default bench command, the reference arm
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02e_pytest.log
( time python bench.py > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err ) 2>&1 | grep real; echo "bench rc=$?"; head -c 300 gpurun_out/r02e_bench.json; echo
( time python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02e_ref.json 2> gpurun_out/r02e_ref.err ) 2>&1 | grep real; head -c 700 gpurun_out/r02e_ref.json; echo
