# round 2, call 65: last check of the committed tree -- build(), smoke(), full GPU suite
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r02g_pytest.log
