"""Summarise an `ncu --page source --csv` dump: top SASS instructions by stall samples, with stall reasons."""
import csv, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
# find header row (starts with Address)
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hi + 1:]:
    if r and r[0] in ("Kernel Name", "Address"):
        break  # next captured launch
    if len(r) == len(hdr):
        data.append(r)
S = col["# Samples"]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[S] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
agg = {s: sum(int(r[col[s]] or 0) for r in data) for s in stalls}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
data.sort(key=lambda r: -int(r[S] or 0))
for r in data[:top]:
    n = int(r[S] or 0)
    st = sorted(((int(r[col[s]] or 0), s) for s in stalls), reverse=True)[:3]
    print(f"{n:7d} {100*n/tot:5.1f}%  exec={r[col['Instructions Executed']]:>10s} thr={r[col['Avg. Threads Executed']]:>5s}  {r[col['Source']][:70]:70s} {[(s[6:],c) for c,s in st if c]}")
