"""Scratch: top SASS instructions by warp-stall samples from `ncu -i rep --page source --csv` output (one kernel)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
tot = 0
for r in rows[2:]:
    if len(r) < len(hdr) or r[col["# Samples"]] == "# Samples" or not r[col["# Samples"]].isdigit():
        continue
    s = int(r[col["# Samples"]] or 0)
    tot += s
    data.append((s, r))
print("total samples", tot)
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for s, r in sorted(data, key=lambda t: -t[0])[:top]:
    st = sorted(((int(r[col[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print(f"{s:7d} {100.0 * s / tot:5.1f}%  {r[col['Source']].strip()[:70]:70s} exec={r[col['Instructions Executed']]:>9s} " +
          " ".join(f"{n}:{v}" for v, n in st if v))
