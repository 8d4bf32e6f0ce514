"""SASS census of every kernel in the built objects: instruction count and the mnemonics that prove the hardware path
(DMMA = FP64 tensor core, UTMALDG = TMA, SYNCS = mbarrier, LDGSTS = cp.async, SHFL, MUFU).  CPU only (cuobjdump)."""
import collections, glob, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
keys = ["DMMA", "UTMALDG", "SYNCS", "LDGSTS", "SHFL", "MUFU", "DFMA", "IMAD", "LDS", "STS", "BAR", "ATOM", "RED"]
print("%-22s %-64s %7s " % ("object", "kernel", "instr") + " ".join("%7s" % k for k in keys))
for obj in sorted(glob.glob(os.path.join(ROOT, "htnorm_b200", "csrc", "_obj", "*.cu.o"))):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    name, cnt = None, None
    out = []
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name: out.append((name, cnt))
            name, cnt = m.group(1), collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and name:
            cnt["_n"] += 1
            cnt[m.group(1)] += 1
    if name: out.append((name, cnt))
    for name, cnt in out:
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"\(anonymous namespace\)::", "", dem)
        dem = re.sub(r"\(.*", "", dem)[:64]
        print("%-22s %-64s %7d " % (os.path.basename(obj), dem, cnt["_n"]) + " ".join("%7d" % cnt[k] for k in keys))
