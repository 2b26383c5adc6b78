"""DRAM traffic and duration of one kernel launch from an `ncu --page raw --csv` export, merged into
profiles/traffic.json (what bench.py reports as roofline.traffic).
python tools/traffic_from_ncu.py <circuit> <kernel> <raw.csv> "<workload>" """
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
circuit, kernel, path, workload = sys.argv[1:5]
rows = list(csv.reader(open(path)))
hdr, units, val = rows[0], rows[1], rows[2]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "second": 1e3,
         "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6}
def get(name):
    i = hdr.index(name)
    return float(val[i].replace(",", "")) * scale[units[i]]
rd, wr, ms = get("dram__bytes_read.sum"), get("dram__bytes_write.sum"), get("gpu__time_duration.sum")
out = os.path.join(ROOT, "profiles", "traffic.json")
t = json.load(open(out)) if os.path.exists(out) else {}
t.setdefault(circuit, {})[kernel] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": rd + wr, "gpu_time_ms": ms}
t[circuit]["workload"] = workload
json.dump(t, open(out, "w"), indent=1)
print(circuit, kernel, rd + wr, ms)
