"""Summarise an `ncu --page source --csv` export: shared-memory wavefronts and stall samples per SASS opcode class.
python tools/ncu_src_summary.py gpurun_out/garble_v6b_src.csv"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; col = {n: i for i, n in enumerate(hdr)}
def f(r, n):
    try: return float(r[col[n]])
    except Exception: return 0.0
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0, 0.0])
tot_w = tot_s = tot_i = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    src = r[col["Source"]].split()
    if not src: continue
    op = src[1] if src[0].startswith("@") else src[0]
    key = op
    if op.startswith("LDS") or op.startswith("STS"):
        key = op + (" [tbl]" if "PRMT" in "" else "")
    a = agg[key]
    a[0] += 1; a[1] += f(r, "Instructions Executed"); a[2] += f(r, "L1 Wavefronts Shared"); a[3] += f(r, "L1 Wavefronts Shared Ideal")
    a[4] += f(r, "# Samples"); a[5] += f(r, "L1 Conflicts Shared N-Way")
    tot_w += f(r, "L1 Wavefronts Shared"); tot_s += f(r, "# Samples"); tot_i += f(r, "Instructions Executed")
print(f"total: warp instr {tot_i:.3e}  shared wavefronts {tot_w:.3e}  samples {tot_s:.0f}")
print(f"{'opcode':28s} {'sites':>5s} {'warp instr':>12s} {'%inst':>6s} {'wavefronts':>12s} {'ideal':>12s} {'%wf':>6s} {'samples%':>8s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{k:28s} {a[0]:5d} {a[1]:12.4e} {100*a[1]/tot_i:6.2f} {a[2]:12.4e} {a[3]:12.4e} {100*a[2]/max(tot_w,1):6.2f} {100*a[4]/max(tot_s,1):8.2f}")
