# round 2, call 40 (second session): sanity of the restored tree (full GPU suite) + the r02 TRAIN per-launch metrics VERDICT asked for
python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest0.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest0.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size
ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02d_train_metrics.csv python tests/prof_train.py > gpurun_out/r02d_ncu_train.log 2>&1
tail -4 gpurun_out/r02d_pytest0.log; tail -2 gpurun_out/r02d_ncu_train.log
