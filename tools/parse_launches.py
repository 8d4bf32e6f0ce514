"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time and share."""
import csv, re, sys, collections
path = sys.argv[1]; last_half = len(sys.argv) > 2 and sys.argv[2] == "half"
lines = [l for l in open(path) if not l.startswith("==")]
rows = []
for row in csv.DictReader(lines):
    if row.get("Metric Name") == "gpu__time_duration.sum":
        v = float(row["Metric Value"].replace(",", ""))
        us = v / 1000 if row["Metric Unit"] in ("ns", "nsecond") else v
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("<unnamed>::", "").replace("void ", "")
        rows.append((int(row["ID"]), name[:60], us, row["Grid Size"]))
if last_half: rows = rows[len(rows) // 2:]
tot = sum(r[2] for r in rows)
if "-v" in sys.argv:
    for r in rows: print(f"{r[0]:4d} {r[2]:9.1f}us {r[3]:>14s} {r[1]}")
agg = collections.OrderedDict()
for _, n, us, g in rows:
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += us
print(f"total {tot:.1f} us over {len(rows)} launches")
for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{us:10.1f} us {100*us/tot:5.1f}%  x{c:<3d} {n}")
