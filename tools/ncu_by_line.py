"""Attribute ncu per-instruction counters to CUDA source lines.
usage: ncu_by_line.py <rep> <object.o> <kernel-substring> [topn]
Joins the SASS rows of `ncu --page source` with `nvdisasm -g` line markers by instruction order."""
import csv, io, os, re, subprocess, sys, tempfile, collections
rep, obj, kname = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
iex, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, capture_output=True)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout.splitlines()
lines = []; cur = None; on = False
for ln in dis:
    if ln.startswith(".text."):
        on = kname in ln
        continue
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln): lines.append(cur)
assert 0 <= len(lines) - len(data) <= 16, (len(lines), len(data))  # nvdisasm also lists the trailing padding NOPs
lines = lines[:len(data)]
ex = collections.Counter(); sa = collections.Counter()
for (loc, r) in zip(lines, data):
    ex[loc] += int(r[iex] or 0); sa[loc] += int(r[isamp] or 0)
tot_e, tot_s = sum(ex.values()), sum(sa.values())
print(f"total executed {tot_e}, samples {tot_s}")
for loc, e in sorted(ex.items(), key=lambda x: -x[1])[:topn]:
    print(f"{str(loc):34s} exec {e:12d} {100*e/tot_e:5.1f}%   samples {sa[loc]:7d} {100*sa[loc]/max(tot_s,1):5.1f}%")
