"""Turn the ncu outputs that gpurun brought back (gpurun_out/) into the tracked summaries under profiles/.

    python profiles/summarize.py r01 gpurun_out/launches.csv gpurun_out/prof_conv_raw.csv
"""
import collections
import csv
import sys

tag, launches, raw = sys.argv[1], sys.argv[2], sys.argv[3]

rows = list(csv.DictReader([l for l in open(launches) if not l.startswith("==")]))
agg = collections.OrderedDict()
for r in rows:
    name = r["Kernel Name"].split("(")[0][:70]
    v = float(r["Metric Value"].replace(",", ""))
    v = v / 1000 if r["Metric Unit"] == "ns" else (v * 1000 if r["Metric Unit"] == "ms" else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
with open(f"profiles/{tag}_launches.md", "w") as f:
    f.write(f"# {tag}: ncu launch list of `python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-train --no-alt` (first {len(rows)} launches)\n\n")
    f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` -- per-launch times are cold-cache and serialised: compare SHARES.\n\n")
    f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {c} | {t:.1f} | {100 * t / tot:.1f}% |\n")

rr = list(csv.reader(open(raw)))
hdr, units = rr[0], rr[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["Grid Size", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor"]
with open(f"profiles/{tag}_conv_ncu.md", "w") as f:
    f.write(f"# {tag}: `ncu --set full --clock-control none -k regex:conv3x3_v2` on the first conv launches of one MC step (`python tests/prof_step.py 10 2`, 10 batched iterations)\n\n")
    f.write("| # | kernel | " + " | ".join(w.split(".")[0] for w in want) + " |\n|---|---|" + "---|" * len(want) + "\n")
    for i, r in enumerate(rr[2:]):
        name = r[idx["Kernel Name"]].split("(")[0].replace("void b2u::", "")
        f.write(f"| {i} | `{name}` | " + " | ".join(f"{r[idx[w]]} {units[idx[w]]}" if w in idx else "-" for w in want) + " |\n")
print("wrote profiles/%s_launches.md and profiles/%s_conv_ncu.md" % (tag, tag))
