"""profiles/traffic.json from the per-launch ncu metrics of one eager MC step:
    python profiles/traffic_from_csv.py gpurun_out/r02d_step_metrics.csv "r02d_s46: <command>" > profiles/traffic.json"""
import collections
import csv
import json
import sys

rows = list(csv.DictReader([l for l in open(sys.argv[1]) if not l.startswith("==")]))
fam = collections.OrderedDict()
seen = set()
for r in rows:
    m = r["Metric Name"]
    if m not in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        continue
    name = r["Kernel Name"].split("(")[0].replace("void ", "").replace("b2u::", "").split("<")[0]
    if name.startswith("at::"):
        continue
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[u]
    f = fam.setdefault(name, {"launches_per_step": 0, "dram_read_mb": 0.0, "dram_write_mb": 0.0})
    if (r["ID"], name) not in seen:
        seen.add((r["ID"], name))
        f["launches_per_step"] += 1
    f["dram_read_mb" if m.endswith("read.sum") else "dram_write_mb"] += v
for f in fam.values():
    f["dram_read_mb"], f["dram_write_mb"] = round(f["dram_read_mb"], 1), round(f["dram_write_mb"], 1)
conv = sum((f["dram_read_mb"] + f["dram_write_mb"]) for k, f in fam.items() if k in ("conv3x3_v2_kernel", "convT_v2_kernel"))
tot = sum((f["dram_read_mb"] + f["dram_write_mb"]) for f in fam.values())
print(json.dumps({"source": sys.argv[2], "iter_batch": 10, "per_family": dict(sorted(fam.items(), key=lambda kv: -(kv[1]["dram_read_mb"] + kv[1]["dram_write_mb"]))),
                  "conv_family_dram_bytes_per_step": conv * 1e6, "step_dram_bytes": tot * 1e6}, indent=1))
