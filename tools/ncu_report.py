"""Text summary of an `ncu --set full --import-source on` report for profiles/: per captured launch the headline raw metrics, the
top warp-stall reasons and the hottest CUDA source lines.  Usage: ncu_report.py report.ncu-rep [kernel-regex] > summary.txt"""
import csv, io, re, subprocess, sys
from collections import defaultdict

rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else None
base = ["ncu", "-i", rep] + (["-k", "regex:" + kre] if kre else [])
raw = subprocess.run(base + ["--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
print(f"# {rep}: {len(rows) - 2} captured launch(es); clock-control none; times are cold-cache, replayed\n")
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")])
    print("==", name)
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"   {k:70s} {r[i]:>16s} {units[i]}")
    st = [(float(r[i].replace(",", "")), h) for i, h in enumerate(hdr) if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and r[i]]
    tot_s = sum(v for v, _ in st) or 1.0
    st = sorted(st, reverse=True)[:7]
    print("   warp-state samples:", ", ".join(f"{h.split('stalled_')[1]} {100 * v / tot_s:.1f}%" for v, h in st))
    print()
src = subprocess.run(base + ["--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
agg = defaultdict(lambda: [0, 0, ""])
fname, h2 = "", None
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": h2 = r; iS = r.index("# Samples"); iI = r.index("Instructions Executed"); continue
    if h2 is None or len(r) != len(h2) or not r[0].isdigit(): continue
    try: s_, n_ = int(r[iS]), int(r[iI])
    except ValueError: continue
    a = agg[(fname, int(r[0]))]; a[0] += s_; a[1] += n_
    if r[1].strip(): a[2] = r[1].strip()[:100]
tot = sum(a[0] for a in agg.values()) or 1
toti = sum(a[1] for a in agg.values()) or 1
print("== hottest source lines over all captured launches (share of attributed stall samples, share of warp-instructions)")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"   {100 * a[0] / tot:5.2f}%s {100 * a[1] / toti:5.2f}%i {f}:{l}: {a[2]}")
