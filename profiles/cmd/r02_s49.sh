# round 2, call 12: r02d evidence set -- full GPU suite, default bench, ncu launch list of the same command, per-launch step
# metrics (one eager MC step, default dtype), --set full capture of the conv kernels
python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; echo "bench rc=$?" >> gpurun_out/r02d_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02d_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-train --no-alt > gpurun_out/r02d_ncu1.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size
ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02d_step_metrics.csv python tests/prof_step.py 10 2 > gpurun_out/r02d_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv3x3_v2 -c 8 -f -o gpurun_out/r02d_conv_full python tests/prof_step.py 10 2 > gpurun_out/r02d_ncu3.log 2>&1
tail -4 gpurun_out/r02d_pytest.log; tail -2 gpurun_out/r02d_bench.err; head -c 400 gpurun_out/r02d_bench.json; tail -2 gpurun_out/r02d_ncu2.log
ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02d_train_metrics.csv python tests/prof_train.py > gpurun_out/r02d_ncu_train.log 2>&1; tail -2 gpurun_out/r02d_ncu_train.log
