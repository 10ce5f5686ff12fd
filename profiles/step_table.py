"""Pivot an `ncu --metrics ... --csv --log-file X.csv` launch list (one row per kernel x metric) into a per-launch
markdown table:  python profiles/step_table.py gpurun_out/s3_step_metrics.csv > profiles/rNN_step_metrics.md"""
import csv
import sys

rows = list(csv.DictReader([l for l in open(sys.argv[1]) if not l.startswith("==")]))
order, data = [], {}
for r in rows:
    k = int(r["ID"])
    if k not in data:
        data[k] = {"name": r["Kernel Name"].split("(")[0].replace("void ", "").replace("b2u::", "")[:44]}
        order.append(k)
    data[k][r["Metric Name"]] = (r["Metric Value"].replace(",", ""), r["Metric Unit"])


def val(d, m, scale=1.0):
    if m not in d:
        return float("nan")
    v, u = d[m]
    v = float(v)
    if u == "ns":
        v /= 1000.0
    elif u == "ms":
        v *= 1000.0
    elif u == "Kbyte":
        v /= 1000.0
    elif u == "Gbyte":
        v *= 1000.0
    elif u == "byte":
        v /= 1e6
    return v * scale


print("| # | kernel | grid | us | DRAM rd MB | DRAM wr MB | DRAM GB/s | dram % | lts % | tensor % | sm % | issue % | regs |")
print("|---|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
tot = {}
for k in order:
    d = data[k]
    us = val(d, "gpu__time_duration.sum")
    rd, wr = val(d, "dram__bytes_read.sum"), val(d, "dram__bytes_write.sum")
    gbs = (rd + wr) / us * 1e3 if us > 0 else 0.0
    print(f"| {k} | `{d['name']}` | {val(d, 'launch__grid_size'):.0f}x{val(d, 'launch__block_size'):.0f} | {us:.1f} | {rd:.1f} | {wr:.1f} | {gbs:.0f} | "
          f"{val(d, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | {val(d, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | "
          f"{val(d, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.0f} | {val(d, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | "
          f"{val(d, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f} | {val(d, 'launch__registers_per_thread'):.0f} |")
    t = tot.setdefault(d["name"], [0, 0.0, 0.0, 0.0])
    t[0] += 1; t[1] += us; t[2] += rd; t[3] += wr
print()
print("| kernel | launches | total us | share | DRAM rd MB | DRAM wr MB | avg GB/s |")
print("|---|---:|---:|---:|---:|---:|---:|")
allus = sum(t[1] for t in tot.values())
for name, t in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name}` | {t[0]} | {t[1]:.1f} | {100 * t[1] / allus:.1f}% | {t[2]:.1f} | {t[3]:.1f} | {(t[2] + t[3]) / t[1] * 1e3:.0f} |")
