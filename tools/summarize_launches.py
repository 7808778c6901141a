"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches, total us, share.
Usage: summarize_launches.py launches.csv [out.csv]"""
import csv, re, sys
from collections import OrderedDict
rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
agg = OrderedDict()
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*$", "", r["Kernel Name"])
    name = re.sub(r"^void ", "", name).replace("(anonymous namespace)::", "").replace("at::", "")
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
out = ["kernel,launches,total_us,avg_us,share"]
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f'"{k}",{n},{us:.1f},{us / n:.2f},{100 * us / tot:.2f}%')
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
