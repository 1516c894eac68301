#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: hottest SASS instructions by stall samples."""
import csv, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[col["# Samples"]] or 0) for r in data)
tot_inst = sum(int(r[col["Instructions Executed"]] or 0) for r in data)
print(f"total samples {tot}, warp instructions executed {tot_inst}, SASS lines {len(data)}")
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[col[h]] or 0) for r in data) for h in stall_cols}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
order = sorted(range(len(data)), key=lambda i: -int(data[i][col["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = data[i]
    st = {h[6:]: int(r[col[h]] or 0) for h in stall_cols if int(r[col[h]] or 0)}
    big = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"{i:5d} {int(r[col['# Samples']]):6d} ex={int(r[col['Instructions Executed']] or 0):9d}  {r[col['Source']][:90]:90s} {big}")
