"""Stall-reason totals of an `ncu --page source --csv` export, plus the top stalled SASS lines.
python tools/ncu_stalls.py file.csv [n_lines]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; col = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot = {n: 0.0 for n in stalls}
lines = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    s = 0.0
    for n in stalls:
        try: v = float(r[col[n]])
        except Exception: v = 0.0
        tot[n] += v; s += v
    lines.append((s, r))
T = sum(tot.values())
print("stall samples:", ", ".join(f"{n[6:]} {100*v/T:.1f}%" for n, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v / T > 0.005))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for s, r in sorted(lines, key=lambda x: -x[0])[:N]:
    top = sorted(((float(r[col[n]] or 0), n[6:]) for n in stalls), reverse=True)[:2]
    print(f"{100*s/T:5.2f}%  {r[col['Address']][-5:]}  {r[col['Source']][:70]:70s} {top[0][1]}:{top[0][0]:.0f} {top[1][1]}:{top[1][0]:.0f}")
