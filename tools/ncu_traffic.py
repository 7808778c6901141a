"""profiles/traffic.json from the ncu --set full captures of the CURRENT kernels (tools/gpu_round2.sh): DRAM bytes (dram__bytes_read.sum +
dram__bytes_write.sum) per launch of env_step_kernel, and summed over the kernels of ONE minibatch of the PPO update.  bench.py reads it for
roofline.traffic.  Usage: ncu_traffic.py <env.ncu-rep> <upd.ncu-rep> [out.json]"""
import csv, io, json, re, subprocess, sys


def rows(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    hdr, units = r[0], r[1]
    out = []
    for x in r[2:]:
        d = {}
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"):
            i = hdr.index(k)
            v = float(x[i].replace(",", ""))
            u = units[i].lower()
            scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}.get(u, 1)
            d[k] = v * scale
        d["name"] = re.sub(r"\(.*", "", x[hdr.index("Kernel Name")]).replace("void ", "").replace("<unnamed>::", "")
        out.append(d)
    return out


env, upd = rows(sys.argv[1]), rows(sys.argv[2])
e = env[0]
# one minibatch = the launches from the first gather_kernel up to (not including) the second one
names = [u["name"] for u in upd]
gi = [i for i, n in enumerate(names) if n.startswith("gather_kernel")]
mb = upd[gi[0]:gi[1]] if len(gi) >= 2 else upd
res = {"env_step_kernel": e["dram__bytes_read.sum"] + e["dram__bytes_write.sum"],
       "env_step_kernel_us_under_ncu": e["gpu__time_duration.sum"],
       "ppo_update_per_minibatch": sum(u["dram__bytes_read.sum"] + u["dram__bytes_write.sum"] for u in mb),
       "ppo_update_kernels": [{"kernel": u["name"], "dram_bytes": u["dram__bytes_read.sum"] + u["dram__bytes_write.sum"], "us_under_ncu": u["gpu__time_duration.sum"]} for u in mb],
       "source": {"env": sys.argv[1], "update": sys.argv[2], "how": "ncu --set full --clock-control none; cold-cache, serialised launches (stepwise update, no graph)"}}
out = sys.argv[3] if len(sys.argv) > 3 else "profiles/traffic.json"
json.dump(res, open(out, "w"), indent=1)
print(json.dumps({k: v for k, v in res.items() if k != "ppo_update_kernels"}, indent=1))
for u in res["ppo_update_kernels"]:
    print(f"  {u['kernel'][:70]:70s} {u['dram_bytes'] / 1e6:8.2f} MB {u['us_under_ncu']:7.2f} us")
