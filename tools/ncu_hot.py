"""Top stall sites from `ncu --page source --csv` (SASS view).  usage: python tools/ncu_hot.py file.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0] != "Address"]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
body_i = list(enumerate(body))
body_i.sort(key=lambda t: -int(t[1][ix["# Samples"]] or 0))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for pos, r in body_i[:n]:
    s = int(r[ix["# Samples"]] or 0)
    top = sorted(((h, int(r[ix[h]] or 0)) for h in stall_cols), key=lambda kv: -kv[1])[:2]
    print(f"{pos:5d} {100.0*s/tot:5.1f}%  {r[ix['Source']][:70]:70s} {top}")
