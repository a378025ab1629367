"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/launch_table.py launches.csv [top_n]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])[:72]
    v = float(r[ix["Metric Value"]])
    unit = r[ix["Metric Unit"]]
    v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
ours = sum(t for k, (n, t) in agg.items() if "pvqa::" in k)
print(f"total {tot/1e3:.3f} ms over {sum(a[0] for a in agg.values())} launches; libpvqa kernels {ours/1e3:.3f} ms "
      f"({100*ours/tot:.1f}%)  [ncu serialises launches with cold caches: compare SHARES, not absolutes]")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{t:9.1f} us {100*t/tot:5.1f}%  n={n:3d}  {k}")
