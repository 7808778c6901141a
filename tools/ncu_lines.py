"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line (samples, instructions).
Usage: ncu_lines.py dump.csv [top]"""
import csv, sys
from collections import defaultdict
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
agg = defaultdict(lambda: [0, 0, ""])
stall = defaultdict(int)
fname = ""
hdr = None
for r in csv.reader(open(sys.argv[1])):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; iS = r.index("# Samples"); iI = r.index("Instructions Executed"); st = [i for i, h in enumerate(r) if h.startswith("stall_") and "Not Issued" not in h]; continue
    if hdr is None or len(r) != len(hdr): continue
    try: s, n = int(r[iS]), int(r[iI])
    except ValueError: continue
    key = (fname, int(r[0]) if r[0].isdigit() else -1)
    a = agg[key]; a[0] += s; a[1] += n
    if r[1].strip(): a[2] = r[1].strip()[:110]
    for i in st:
        try: stall[hdr[i]] += int(r[i])
        except ValueError: pass
tot = sum(a[0] for a in agg.values()); toti = sum(a[1] for a in agg.values())
print(f"samples {tot}  warp-instructions {toti}")
print("stalls:", ", ".join(f"{k[6:]} {100*v/max(1,sum(stall.values())):.1f}%" for k, v in sorted(stall.items(), key=lambda kv: -kv[1])[:8]))
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*a[0]/tot:5.2f}%s {100*a[1]/toti:5.2f}%i {f}:{l}: {a[2]}")
