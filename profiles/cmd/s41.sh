for nb in 16 10 12 20 10 16 12 8; do
python bench.py --steps 40 --no-cpu --no-e2e --no-train --iter-batch $nb > gpurun_out/s41_nb$nb.json 2> gpurun_out/s41_nb$nb.err
python - <<P
import json
d=json.load(open('gpurun_out/s41_nb$nb.json')); print($nb, round(d['value'],1), round(d['ms_per_step'],3), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])
P
done
