"""Summarise an ncu --page source --csv dump: top SASS instructions by stall samples."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
tot = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        s = int(r[ix["# Samples"]])
    except ValueError:
        continue
    tot += s
    st = {c: int(r[ix[c]] or 0) for c in stall_cols}
    data.append((s, r[ix["Source"]].strip(), st, r[ix["Instructions Executed"]]))
print("total samples", tot, "instructions", len(data))
agg = {}
for s, src, st, _ in data:
    for c, v in st.items():
        agg[c] = agg.get(c, 0) + v
print("stall totals:", sorted(((v, c) for c, v in agg.items() if v), reverse=True)[:8])
for i, (s, src, st, ne) in enumerate(data):
    pass
order = sorted(range(len(data)), key=lambda i: -data[i][0])[:top]
for i in sorted(order):
    s, src, st, ne = data[i]
    main = sorted(((v, c) for c, v in st.items() if v), reverse=True)[:2]
    print(f"{i:5d} {s:7d} {100*s/tot:5.1f}% exec={ne:>9} {src[:70]:70s} {main}")
