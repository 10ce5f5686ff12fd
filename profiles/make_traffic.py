"""profiles/traffic.json from an ncu per-launch metrics CSV of ONE Monte-Carlo step (tests/prof_step.py NB 2):

    python profiles/make_traffic.py gpurun_out/s12_step_metrics.csv 10 r01_s12 > profiles/traffic.json

DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) summed per kernel family over the step's launches."""
import csv
import json
import sys

path, nb, tag = sys.argv[1], int(sys.argv[2]), sys.argv[3]
rows = list(csv.DictReader([l for l in open(path) if not l.startswith("==")]))
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
fam = {}
for r in rows:
    if r["Metric Name"] not in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        continue
    name = r["Kernel Name"].split("(")[0].replace("void ", "").replace("b2u::", "").split("<")[0]
    f = fam.setdefault(name, {"launches": set(), "read": 0.0, "write": 0.0})
    f["launches"].add(r["ID"])
    f["read" if "read" in r["Metric Name"] else "write"] += float(r["Metric Value"].replace(",", "")) * scale[r["Metric Unit"]]
out = {"source": f"{tag}: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, python tests/prof_step.py {nb} 2",
       "iter_batch": nb, "per_family": {}}
conv = 0.0
step_total = 0.0
for name, f in sorted(fam.items(), key=lambda kv: -(kv[1]["read"] + kv[1]["write"])):
    tot = f["read"] + f["write"]
    out["per_family"][name] = {"launches_per_step": len(f["launches"]), "dram_read_mb": round(f["read"] / 1e6, 1),
                               "dram_write_mb": round(f["write"] / 1e6, 1)}
    step_total += tot
    if name.startswith("conv3x3_v2") or name.startswith("convT_v2"):
        conv += tot
out["conv_family_dram_bytes_per_step"] = conv
out["step_dram_bytes"] = step_total
print(json.dumps(out, indent=1))
