python -m pytest tests -m gpu -x -q > gpurun_out/s36_pytest.log 2>&1; tail -3 gpurun_out/s36_pytest.log
python bench.py > gpurun_out/s36_bench.json 2> gpurun_out/s36_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s36_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-train > gpurun_out/s36_ncu1.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size
ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/s36_step_metrics.csv python tests/prof_step.py 10 2 > gpurun_out/s36_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv3x3_v2 -c 8 -f -o gpurun_out/s36_conv_full python tests/prof_step.py 10 2 > gpurun_out/s36_ncu3.log 2>&1
cut -c1-200 gpurun_out/s36_bench.json
