"""Per-kernel counts of the SASS mnemonics that prove Blackwell-native code (B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld,
UTMALDG / UTMASTG / UTMAREDG = TMA tensor load / store / reduce, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier.
Usage: python tools/sass_evidence.py [libgrx_b200.so] > profiles/sass_evidence.txt   (needs cuobjdump; no GPU)"""
import collections, os, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "wiki-grx-gym_b200", "libgrx_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
keys = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "SYNCS", "HMMA", "SHFL", "REDG", "NANOSLEEP"]
pat = re.compile(r"\b(" + "|".join(keys) + r")\b")
print("# " + os.path.basename(so) + ": SASS mnemonic counts per kernel (static)")
print("kernel,instructions," + ",".join(keys))
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    ops = collections.Counter(m.group(1) for m in pat.finditer(f))
    n = len(re.findall(r"/\*[0-9a-f]{4,5}\*/", f))
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    dem = re.sub(r"\(anonymous namespace\)::", "", dem)
    dem = re.sub(r"\(.*", "", dem)
    print(f'"{dem}",{n},' + ",".join(str(ops.get(k, 0)) for k in keys))
