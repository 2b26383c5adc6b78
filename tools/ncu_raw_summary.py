"""Key figures of `ncu --page raw --csv` exports as one markdown table row per file.
python tools/ncu_raw_summary.py label=file.csv [label=file.csv ...]"""
import csv, sys
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
COLS = [("time ms", "gpu__time_duration.sum", 1), ("regs", "launch__registers_per_thread", 0),
        ("warps/SM", "sm__warps_active.avg.per_cycle_active", 0),
        ("LSU pipe %", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", 0),
        ("smem wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", 0),
        ("bank conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 0),
        ("issue %", "sm__issue_active.avg.pct_of_peak_sustained_elapsed", 0),
        ("ALU pipe %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 0),
        ("FMA pipe %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 0),
        ("warp inst", "smsp__inst_executed.sum", 0),
        ("DRAM rd GB", "dram__bytes_read.sum", 2), ("DRAM wr GB", "dram__bytes_write.sum", 2),
        ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0)]
print("| kernel | " + " | ".join(c[0] for c in COLS) + " |")
print("|---|" + "---|" * len(COLS))
for arg in sys.argv[1:]:
    label, path = arg.split("=", 1)
    rows = list(csv.reader(open(path)))
    hdr, units, val = rows[0], rows[1], rows[2]
    out = []
    for name, metric, kind in COLS:
        if metric not in hdr:
            out.append("-"); continue
        i = hdr.index(metric)
        try:
            v = float(val[i].replace(",", ""))
        except ValueError:
            out.append(val[i]); continue
        if kind == 1: v *= SCALE.get(units[i], 1.0)
        if kind == 2: v *= SCALE.get(units[i], 1.0) / 1e9
        out.append(f"{v:.3e}" if v >= 1e6 else f"{v:.2f}" if v != int(v) else f"{int(v)}")
    print(f"| {label} | " + " | ".join(out) + " |")
