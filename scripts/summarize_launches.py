"""Summarise an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv) of
bench.py: per kernel launches, total device time, share of the captured step, average DRAM bytes per launch.
Writes a markdown table and profiles/traffic.json (read by bench.py for roofline.traffic).

python scripts/summarize_launches.py launches.csv out.md [traffic.json]"""
import collections
import csv
import json
import re
import sys

src, out_md = sys.argv[1], sys.argv[2]
traffic_path = sys.argv[3] if len(sys.argv) > 3 else None
rows = [r for r in csv.reader(open(src, errors="ignore")) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4]
    m = re.search(r"(\w+_kernel)", name)
    short = m.group(1) if m else name[:40]
    metric, unit, val = r[-3], r[-2], float(r[-1].replace(",", ""))
    a = agg.setdefault(short, {"ids": set(), "ns": 0.0, "rd": 0.0, "wr": 0.0})
    a["ids"].add(r[0])
    scale = {"ns": 1.0, "us": 1e3, "ms": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    if metric.startswith("gpu__time_duration"):
        a["ns"] += val * scale
    elif metric.startswith("dram__bytes_read"):
        a["rd"] += val * scale
    elif metric.startswith("dram__bytes_write"):
        a["wr"] += val * scale
total = sum(a["ns"] for a in agg.values())
lines = [f"ncu launch list `{src.split('/')[-1]}`: {sum(len(a['ids']) for a in agg.values())} launches, {total / 1e6:.2f} ms of device time "
         "(cold-cache, serialised: compare SHARES with bench.py's CUDA-event shares, not absolutes)", "",
         "| kernel | launches | total ms | share | avg us | avg DRAM MB / launch (rd + wr) |", "|---|---|---|---|---|---|"]
traffic = {"source": src.split("/")[-1]}
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
    n = len(a["ids"])
    lines.append(f"| `{k}` | {n} | {a['ns'] / 1e6:.2f} | {a['ns'] / total:.3f} | {a['ns'] / n / 1e3:.1f} | {a['rd'] / n / 1e6:.1f} + {a['wr'] / n / 1e6:.1f} |")
    traffic[k] = {"launches": n, "dram_bytes_per_launch": (a["rd"] + a["wr"]) / n, "avg_us": a["ns"] / n / 1e3, "share": a["ns"] / total}
open(out_md, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
if traffic_path:
    json.dump(traffic, open(traffic_path, "w"), indent=1)
