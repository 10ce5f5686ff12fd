# round 2, call 46: r02d per-launch metrics of ONE eager MC step (new defaults: TMA-store epilogues, levels 1..4 fused) + timeline
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size
ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02d_step_metrics.csv python tests/prof_step.py 10 2 > gpurun_out/r02d_ncu2.log 2>&1
tail -2 gpurun_out/r02d_ncu2.log
timeout 300 python tests/exp_timeline.py > gpurun_out/r02d_timeline.log 2>&1; tail -5 gpurun_out/r02d_timeline.log
