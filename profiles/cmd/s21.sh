for nb in 10 12 14 16 20; do
python bench.py --steps 30 --no-cpu --no-e2e --no-train --iter-batch $nb > gpurun_out/s21_nb$nb.json 2> gpurun_out/s21_nb$nb.err
python - <<P
import json
d=json.load(open('gpurun_out/s21_nb$nb.json')); print($nb, round(d['value'],1), round(d['ms_per_step'],3), round(d['roofline']['frac'],4), round(d['roofline']['conv_ms_per_step'],3), d['clocks']['sm_mhz'])
P
done
