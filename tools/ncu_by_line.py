"""Attribute the warp instructions / stall samples of an `ncu --page source --csv` export to CUDA source lines, using
`nvdisasm -g -c` output of the same cubin (line info from -lineinfo).
  cuobjdump -xelf all mpc_b200/libgcb200.so; nvdisasm -g -c gc_eval.sm_100a.cubin > eval_all.sass
  python tools/ncu_by_line.py eval_all.sass '<mangled kernel name>' src.csv [top_n]"""
import collections, csv, re, sys
sass, fun, src = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
line_of = {}
cur = None
inside = False
for l in open(sass):
    if l.startswith(".text."):
        inside = l.strip() == f".text.{fun}:"
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
col = {n: i for i, n in enumerate(rows[hi])}
base = int(rows[hi + 1][0], 16)
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
ti = ts = 0.0
for r in rows[hi + 1:]:
    if len(r) < len(rows[hi]):
        continue
    off = int(r[0], 16) - base
    key = line_of.get(off, ("?", 0))
    def f(n):
        try: return float(r[col[n]])
        except Exception: return 0.0
    a = agg[key]
    a[0] += f("Instructions Executed"); a[1] += f("# Samples"); a[2] += f("L1 Wavefronts Shared")
    ti += f("Instructions Executed"); ts += f("# Samples")
print(f"total warp instr {ti:.3e} samples {ts:.0f}; mapped lines {len(agg)}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]:22s} {k[1]:5d}  inst {100*a[0]/ti:5.2f}%  samples {100*a[1]/ts:5.2f}%  smem wf {a[2]:.3e}")
