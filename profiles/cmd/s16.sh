python -m pytest tests -m gpu -x -q > gpurun_out/s16_pytest.log 2>&1; tail -3 gpurun_out/s16_pytest.log
python bench.py --steps 30 --no-cpu --no-e2e > gpurun_out/s16_bench.json 2> gpurun_out/s16_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/s16_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['other_kernels_ms_per_step'], d['train']['value'], d['train']['ms_per_step'])
P
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size
ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/s16_train_metrics.csv python tests/prof_train.py > gpurun_out/s16_ncu_train.log 2>&1
tail -2 gpurun_out/s16_ncu_train.log
